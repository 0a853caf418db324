// placeholder until the BDPT wavefront lands (next commit)
#include <string>
#include "wavefront.h"
namespace wrt { void bdpt_destroy(wrt_wavefront*) {} }
extern "C" {
int wrt_render_bdpt(wrt_scene*, const wrt_camera*, const wrt_bdpt_params*, float*) { wrt::set_error("wrt_render_bdpt: not built yet"); return WRT_ERR_INVALID; }
int wrt_render_bdpt_dev(wrt_scene*, const wrt_camera*, const wrt_bdpt_params*, float*, void*) { wrt::set_error("wrt_render_bdpt: not built yet"); return WRT_ERR_INVALID; }
}
