/* aq_bvh_build.h — host BVH8 builder interface (see aq_bvh_build.cpp). */
#ifndef AQ_BVH_BUILD_H
#define AQ_BVH_BUILD_H
#include <vector>

#include "aq_bvh.h"

struct aq_bvh8 {
    std::vector<aq_u4> nodes; /* AQ_NODE_WORDS per node, node 0 = root */
    std::vector<aq_f4> tris;  /* AQ_TRI_WORDS per triangle record */
    uint32_t max_depth;
    uint32_t n_bvh2_nodes;
    float sah_cost;
};

/* returns 0 on success, -1 if the tree is deeper than the traversal stack allows */
int aq_build_bvh8(const float* positions, const uint32_t* indices, uint32_t n_tris, int n_threads,
                  aq_bvh8* out);

#if defined(__CUDACC__)
#include <cuda_runtime.h>
/* device builder (aq_bvh_build_gpu.cu): Morton order -> binary tree (tree_mode 0: PLOC for surface-like input, the
 * radix tree for a soup; 1: radix tree; 2: PLOC) -> cost-optimal collapse -> BVH8; 0 ok, -1 too deep / overflow,
 * -2 CUDA error */
int aq_build_bvh8_device(cudaStream_t st, const float* d_pos, const uint32_t* d_idx, uint32_t n_tris,
                         aq_u4** d_nodes, size_t* n_node_words, aq_f4** d_tris, uint32_t* max_depth,
                         cudaError_t* err, int tree_mode);
#endif
#endif
