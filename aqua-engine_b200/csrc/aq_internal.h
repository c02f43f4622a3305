/* aq_internal.h — hooks shared between translation units of libaqua_cuda.so (not exported
 * through include/aqua_cuda.h). */
#ifndef AQ_INTERNAL_H
#define AQ_INTERNAL_H
#include <cuda_runtime.h>

#include "aqua_cuda.h"

/* device film of the last aq_render_device_async(scene, cfg, NULL) */
void* aq_internal_film(aq_scene* s);
cudaStream_t aq_internal_stream(aq_scene* s);
int aq_internal_device(aq_scene* s);
/* copy src's built BVH8 to dst (another device) instead of rebuilding it on the host */
int aq_internal_clone_accel(aq_scene* dst, aq_scene* src);
int aq_internal_set_error(aq_ctx* c, int code, const char* msg);
cudaStream_t aq_internal_ctx_stream(aq_ctx* c);
int aq_internal_ctx_device(aq_ctx* c);
#endif
