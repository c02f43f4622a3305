/* aq_multi.cu — placeholder, replaced below by the NCCL implementation */
#include "aqua_cuda.h"
extern "C" int aq_render_multi(const aq_scene_desc*, const aq_integrator_cfg*, int, const int*, float*, aq_stats*) { return AQ_ERR_UNSUPPORTED; }
