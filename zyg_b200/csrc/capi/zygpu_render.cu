// placeholder until the device passes land
#include "../../../include/zygpu.h"
extern "C" {
int zygpu_upload_scene(zygpu_device*, const ZygpuScene*) { return -1; }
int zygpu_set_view(zygpu_device*, const ZygpuView*) { return -1; }
int zygpu_clear_film(zygpu_device*) { return -1; }
int zygpu_render(zygpu_device*, uint32_t, uint32_t) { return -1; }
int zygpu_resolve(zygpu_device*, float*, uint32_t) { return -1; }
int zygpu_download_film(zygpu_device*, float*, uint32_t) { return -1; }
int zygpu_upload_film(zygpu_device*, const float*, uint32_t) { return -1; }
void* zygpu_film_device(zygpu_device*, uint64_t*) { return nullptr; }
int zygpu_synchronize(zygpu_device*) { return -1; }
int zygpu_render_stats(zygpu_device*, ZygpuRenderStats*) { return -1; }
}
