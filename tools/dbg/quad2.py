import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, 'tests')
import numpy as np
import oracle_lib as oracle
from zyg_b200 import lib
import test_trace_gpu as T
orig = T.check_closest
def dbg(dev, mesh, rays, want_hits=0):
    nodes, tris, pos, original = T.arrays(mesh)
    mid = dev.upload_mesh(mesh)
    ref = oracle.trace_closest(nodes, tris, pos, rays)
    exact = dev.trace_batch(mid, lib.CLOSEST_BINARY, rays)
    bad = np.nonzero((exact.view(np.uint32).reshape(-1,4) != ref.view(np.uint32).reshape(-1,4)).any(1))[0]
    print('mid', mid, 'bad', len(bad), bad[:20])
    for i in bad[:6]:
        print(i, rays[i], 'ref', ref[i], ref[i:i+1].view(np.uint32), 'dev', exact[i], exact[i:i+1].view(np.uint32))
    return orig(dev, mesh, rays, want_hits)
T.check_closest = dbg
dev = lib.Device(0)
for s in sys.argv[1:]:
    try:
        T.test_small_meshes_and_ties(dev, s)
        print(s, 'ok')
    except AssertionError as e:
        print(s, 'FAILED')
