/*
 * EXPERIMENTAL — NOT PART OF THE PRODUCT BUILD, NEVER RUN ON A GPU YET (written at the end of
 * round 1, when the round's GPU budget was spent).  Starting point for round 2.
 *
 * The nrc integrator's query MLP (aq_nrc.h: 64 -> 4 x 64 ReLU -> 3) on the 5th-generation tensor
 * cores: one CTA of 128 threads owns a tile of 128 queries; per layer ONE elected thread issues
 * four tcgen05.mma (M = 128, N = 64, K = 16, bf16 x bf16 -> fp32) whose accumulator lives in
 * TMEM (64 columns); the four warps read their 32 TMEM lanes back with tcgen05.ld, apply the
 * ReLU, round to bf16 and write the next layer's A operand straight into shared memory in the
 * K-major core-matrix layout the MMA reads.  The five weight matrices stay resident in shared
 * memory (40 KB, written once per CTA).  No TMA (the operands are produced by the CTA itself), no
 * cluster, cta_group::1.
 *
 * This trades the bit-exact parity of the fp32 query kernel (aq_k_nrc_query) for a tolerance:
 * operands are rounded to bf16, the accumulation order inside the MMA is unspecified.  The
 * harness below therefore compares against a CPU evaluation with the same bf16 roundings and
 * reports the largest deviation.
 *
 * build:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o nrc_tc_probe tools/experimental/nrc_tcgen05_query.cu
 * run:    timeout 20 ./nrc_tc_probe        (every mbarrier wait is bounded: a lost arrive traps, it cannot hang)
 *
 * Layout of an operand tile with R rows (M or N) and 64 columns (K), bf16, no swizzle:
 *   core matrix = 8 rows x 8 elements (16 B per row, rows contiguous: 128 B)
 *   element (r, k) at byte  (r/8)*128 + (k/8)*(R/8)*128 + (r%8)*16 + (k%8)*2
 *   => matrix descriptor: stride-dimension byte offset (next 8-row group) = 128,
 *      leading-dimension byte offset (next core matrix along K) = (R/8)*128,
 *      one K = 16 MMA step advances the start address by 2 * LBO.
 */
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>

#define NQ 128          /* queries per tile = UMMA M = TMEM lanes */
#define WID 64          /* K of every layer, N of the hidden layers */
#define NOUT 16         /* N of the output layer (3 used) */
#define NHID 4

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

/* 64-bit shared-memory matrix descriptor, SWIZZLE_NONE, K-major (cute/arch/mma_sm100_desc.hpp) */
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);             /* start address, bits [0,14) */
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;    /* leading dimension byte offset, bits [16,30) */
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;    /* stride dimension byte offset, bits [32,46) */
    d |= (uint64_t)1 << 46;                               /* descriptor version 1 (sm_100) */
    return d;                                             /* base offset 0, lbo mode 0, layout type 0 = no swizzle */
}

/* 32-bit instruction descriptor: D = fp32, A = B = bf16, both K-major, dense */
__host__ __device__ constexpr uint32_t make_idesc(uint32_t M, uint32_t N) {
    return (1u << 4)            /* c_format  = F32  */
         | (1u << 7)            /* a_format  = BF16 */
         | (1u << 10)           /* b_format  = BF16 */
         | ((N >> 3) << 17)     /* n_dim */
         | ((M >> 4) << 24);    /* m_dim */
}

__device__ __forceinline__ void mma_bf16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
        "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
        : "memory");
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
/* bounded wait: a lost arrive traps instead of hanging the GPU */
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done = 0;
    for (uint32_t spin = 0; spin < (1u << 24); ++spin) {
        asm volatile(
            "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
        if (done) return;
    }
    asm volatile("trap;\n");
}

/* byte offset of element (r, k) in a tile of R rows (see the header) */
__device__ __host__ __forceinline__ uint32_t tile_off(uint32_t r, uint32_t k, uint32_t R) {
    return (r >> 3) * 128u + (k >> 3) * (R >> 3) * 128u + (r & 7u) * 16u + (k & 7u) * 2u;
}

/* wt: the five weight matrices already in operand layout (host-prepared): hidden layer l at byte
 * l*8192 as a [N=64 rows][K=64] tile with B(n,k) = W_l[k][n]; output layer at 4*8192 as a [16][64]
 * tile (rows 3..15 zero).  x: [n][64] fp32 features, y: [n][3] fp32. */
__global__ void __launch_bounds__(NQ) nrc_query_tc(const uint8_t* __restrict__ wt, const float* __restrict__ x,
                                                   float* __restrict__ y, uint32_t n) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* sA = smem;                          /* 128 x 64 bf16 = 16 KB */
    uint8_t* sW = smem + NQ * WID * 2;            /* 4 * 8 KB + 2 KB */
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const uint32_t tid = threadIdx.x, warp = tid >> 5;

    for (uint32_t k = tid; k < (NHID * WID * WID * 2 + NOUT * WID * 2) / 16; k += NQ)
        reinterpret_cast<uint4*>(sW)[k] = reinterpret_cast<const uint4*>(wt)[k];
    if (tid == 0) mbar_init(&bar, 1);
    if (warp == 0) { /* one warp allocates 64 TMEM columns and gives the permit back */
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;\n" ::"r"(smem_u32(&tmem_base_s)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const uint32_t tmem = tmem_base_s;
    uint32_t phase = 0;

    for (uint32_t base = blockIdx.x * NQ; base < n; base += gridDim.x * NQ) {
        /* ---- this thread's query row -> bf16 A operand (row tid) */
        const uint32_t q = base + tid;
        for (uint32_t kc = 0; kc < WID / 8; ++kc) {
            __nv_bfloat162 v[4];
            for (int e = 0; e < 4; ++e) {
                float a = q < n ? x[(size_t)q * WID + kc * 8 + 2 * e] : 0.f;
                float b = q < n ? x[(size_t)q * WID + kc * 8 + 2 * e + 1] : 0.f;
                v[e] = __floats2bfloat162_rn(a, b);
            }
            *reinterpret_cast<uint4*>(sA + tile_off(tid, kc * 8, NQ)) = *reinterpret_cast<uint4*>(v);
        }
        for (uint32_t l = 0; l <= NHID; ++l) {
            /* generic-proxy writes of sA -> visible to the tensor core (async proxy), all rows written */
            asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
            __syncthreads();
            const uint32_t N = l < NHID ? WID : NOUT;
            if (tid == 0) {
                asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
                const uint32_t a0 = smem_u32(sA), b0 = smem_u32(sW + l * WID * WID * 2);
                const uint32_t lbo_a = (NQ / 8) * 128, lbo_b = (N / 8) * 128;
                const uint32_t idesc = make_idesc(NQ, N);
                for (uint32_t s = 0; s < WID / 16; ++s)
                    mma_bf16(tmem, make_desc(a0 + s * 2 * lbo_a, lbo_a, 128), make_desc(b0 + s * 2 * lbo_b, lbo_b, 128),
                             idesc, s > 0);
                /* commit: arrives on the mbarrier when the MMAs above have completed */
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(&bar)) : "memory");
            }
            mbar_wait(&bar, phase);
            phase ^= 1u;
            asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
            /* ---- accumulator row = TMEM lane (32*warp + lane): read it back */
            const uint32_t taddr = tmem + ((warp * 32u) << 16);
            if (l < NHID) {
                uint32_t r[64];
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x64.b32 "
                    "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                    "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, "
                    "%32, %33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, %46, %47, "
                    "%48, %49, %50, %51, %52, %53, %54, %55, %56, %57, %58, %59, %60, %61, %62, %63}, [%64];\n"
                    : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                      "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                      "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                      "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]),
                      "=r"(r[32]), "=r"(r[33]), "=r"(r[34]), "=r"(r[35]), "=r"(r[36]), "=r"(r[37]), "=r"(r[38]), "=r"(r[39]),
                      "=r"(r[40]), "=r"(r[41]), "=r"(r[42]), "=r"(r[43]), "=r"(r[44]), "=r"(r[45]), "=r"(r[46]), "=r"(r[47]),
                      "=r"(r[48]), "=r"(r[49]), "=r"(r[50]), "=r"(r[51]), "=r"(r[52]), "=r"(r[53]), "=r"(r[54]), "=r"(r[55]),
                      "=r"(r[56]), "=r"(r[57]), "=r"(r[58]), "=r"(r[59]), "=r"(r[60]), "=r"(r[61]), "=r"(r[62]), "=r"(r[63])
                    : "r"(taddr)
                    : "memory");
                asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
                /* ReLU, round to bf16, write the next layer's A row */
                for (uint32_t kc = 0; kc < WID / 8; ++kc) {
                    __nv_bfloat162 v[4];
                    for (int e = 0; e < 4; ++e)
                        v[e] = __floats2bfloat162_rn(fmaxf(__uint_as_float(r[kc * 8 + 2 * e]), 0.f),
                                                     fmaxf(__uint_as_float(r[kc * 8 + 2 * e + 1]), 0.f));
                    *reinterpret_cast<uint4*>(sA + tile_off(tid, kc * 8, NQ)) = *reinterpret_cast<uint4*>(v);
                }
            } else {
                uint32_t r[4];
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];\n"
                             : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                             : "r"(taddr)
                             : "memory");
                asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
                if (q < n) {
                    y[(size_t)q * 3 + 0] = __uint_as_float(r[0]);
                    y[(size_t)q * 3 + 1] = __uint_as_float(r[1]);
                    y[(size_t)q * 3 + 2] = __uint_as_float(r[2]);
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;\n" ::"r"(tmem) : "memory");
}

/* ------------------------------------------------------------------ harness */
static float bf16r(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }

int main() {
    const uint32_t n = 1u << 16;
    std::vector<float> W((NHID * WID * WID) + WID * 4), x((size_t)n * WID), y((size_t)n * 3), ref((size_t)n * 3);
    srand(1);
    for (auto& w : W) w = ((rand() % 2001) - 1000) * 2.1e-4f;
    for (auto& v : x) v = (rand() % 1000) * 1e-3f;
    /* operand layout of the weights: B(n,k) = W_l[k][n] (W_l input-major as in aq_nrc.h) */
    std::vector<uint8_t> wt(NHID * WID * WID * 2 + NOUT * WID * 2, 0);
    auto put = [&](size_t base, uint32_t R, uint32_t r, uint32_t k, float v) {
        __nv_bfloat16 b = __float2bfloat16_rn(v);
        *reinterpret_cast<__nv_bfloat16*>(&wt[base + tile_off(r, k, R)]) = b;
    };
    for (uint32_t l = 0; l < NHID; ++l)
        for (uint32_t k = 0; k < WID; ++k)
            for (uint32_t j = 0; j < WID; ++j) put((size_t)l * WID * WID * 2, WID, j, k, W[l * WID * WID + k * WID + j]);
    for (uint32_t k = 0; k < WID; ++k)
        for (uint32_t c = 0; c < 3; ++c) put((size_t)NHID * WID * WID * 2, NOUT, c, k, W[NHID * WID * WID + k * 4 + c]);
    /* CPU reference with the same roundings (fp32 accumulate, ascending k) */
    for (uint32_t q = 0; q < n; ++q) {
        float a[WID], b[WID];
        for (int k = 0; k < WID; ++k) a[k] = bf16r(x[(size_t)q * WID + k]);
        for (uint32_t l = 0; l < NHID; ++l) {
            for (int j = 0; j < WID; ++j) {
                float acc = 0.f;
                for (int k = 0; k < WID; ++k) acc += a[k] * bf16r(W[l * WID * WID + k * WID + j]);
                b[j] = bf16r(fmaxf(acc, 0.f));
            }
            for (int k = 0; k < WID; ++k) a[k] = b[k];
        }
        for (int c = 0; c < 3; ++c) {
            float acc = 0.f;
            for (int k = 0; k < WID; ++k) acc += a[k] * bf16r(W[NHID * WID * WID + k * 4 + c]);
            ref[(size_t)q * 3 + c] = acc;
        }
    }
    uint8_t* d_wt;
    float *d_x, *d_y;
    cudaMalloc(&d_wt, wt.size());
    cudaMalloc(&d_x, x.size() * 4);
    cudaMalloc(&d_y, y.size() * 4);
    cudaMemcpy(d_wt, wt.data(), wt.size(), cudaMemcpyHostToDevice);
    cudaMemcpy(d_x, x.data(), x.size() * 4, cudaMemcpyHostToDevice);
    cudaMemset(d_y, 0, y.size() * 4);
    const size_t smem = NQ * WID * 2 + NHID * WID * WID * 2 + NOUT * WID * 2;
    cudaFuncSetAttribute(nrc_query_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    nrc_query_tc<<<148 * 2, NQ, smem>>>(d_wt, d_x, d_y, n);
    cudaError_t err = cudaDeviceSynchronize();
    if (err != cudaSuccess) {
        printf("kernel failed: %s\n", cudaGetErrorString(err));
        return 2;
    }
    cudaEventRecord(e0);
    for (int i = 0; i < 10; ++i) nrc_query_tc<<<148 * 2, NQ, smem>>>(d_wt, d_x, d_y, n);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaMemcpy(y.data(), d_y, y.size() * 4, cudaMemcpyDeviceToHost);
    double maxerr = 0, maxref = 0;
    for (size_t k = 0; k < y.size(); ++k) {
        maxerr = fmax(maxerr, fabs((double)y[k] - ref[k]));
        maxref = fmax(maxref, fabs((double)ref[k]));
    }
    printf("n=%u  %.3f ms/launch  %.2f TFLOP/s  max |err| = %.3g (max |ref| = %.3g)\n", n, ms / 10,
           2.0 * n * (NHID * WID * WID + WID * 3) / (ms / 10 * 1e-3) / 1e12, maxerr, maxref);
    return maxerr <= 2e-2 * fmax(maxref, 1e-3) ? 0 : 1;
}
