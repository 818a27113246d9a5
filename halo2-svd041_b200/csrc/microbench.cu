// Integer-pipe micro-benchmarks: the measured denominators of the mat-mul roofline (DESIGN.md "K1").
// Every kernel runs `iters` rounds of fully unrolled, register-resident work on all SMs with enough
// independent streams per thread to cover the pipe latency, and keeps the result live.
#include "common.cuh"
#include "fr_acc.cuh"

namespace h2svd {

namespace {

constexpr int MB_THREADS = 256;
constexpr int MB_ILP = 8;

// kind 0: mad.lo.u32 (SASS IMAD), 8 independent chains per thread
__global__ void __launch_bounds__(MB_THREADS) mb_imad_lo(uint32_t* out, int iters, uint32_t seed) {
    uint32_t x[MB_ILP];
#pragma unroll
    for (int i = 0; i < MB_ILP; i++) x[i] = seed + threadIdx.x * 7 + i;
    const uint32_t m = seed | 1u;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 16; r++)
#pragma unroll
            for (int i = 0; i < MB_ILP; i++) x[i] = x[i] * m + x[(i + 1) % MB_ILP];
    }
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < MB_ILP; i++) s ^= x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// kind 1: mad.wide.u32 (SASS IMAD.WIDE.U32), 8 independent 64-bit accumulators per thread
__global__ void __launch_bounds__(MB_THREADS) mb_imad_wide(uint64_t* out, int iters, uint32_t seed) {
    uint64_t x[MB_ILP];
#pragma unroll
    for (int i = 0; i < MB_ILP; i++) x[i] = seed + threadIdx.x * 7 + i;
    uint32_t m = seed | 1u;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 16; r++)
#pragma unroll
            for (int i = 0; i < MB_ILP; i++)
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(x[i]) : "r"((uint32_t)x[(i + 1) % MB_ILP]), "r"(m));
    }
    uint64_t s = 0;
#pragma unroll
    for (int i = 0; i < MB_ILP; i++) s ^= x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// kind 2: the mat-mul's carry chains (4 x IMAD.WIDE.U32[.X] + 1 x IADD3.X each), 4 chains in flight
__global__ void __launch_bounds__(MB_THREADS) mb_chain(uint64_t* out, int iters, uint32_t seed) {
    uint64_t d[4][4];
    uint32_t cnt[4] = {0, 0, 0, 0};
#pragma unroll
    for (int c = 0; c < 4; c++)
#pragma unroll
        for (int i = 0; i < 4; i++) d[c][i] = seed + threadIdx.x + c * 4 + i;
    const uint32_t a0 = seed | 1u, a1 = seed * 3u, a2 = seed * 5u, a3 = seed * 7u;
    uint32_t b = seed ^ threadIdx.x;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 8; r++) {
#pragma unroll
            for (int c = 0; c < 4; c++) fr::chain4(d[c], cnt[c], a0, a1, a2, a3, b);
            b += 0x9e3779b9u;
        }
    }
    uint64_t s = 0;
#pragma unroll
    for (int c = 0; c < 4; c++) {
        s ^= cnt[c];
#pragma unroll
        for (int i = 0; i < 4; i++) s ^= d[c][i];
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// kind 3: complete lazy 8x8 multiply-accumulates, two accumulators per thread, operands in registers
__global__ void __launch_bounds__(MB_THREADS) mb_mulacc(uint64_t* out, int iters, uint32_t seed) {
    fr::WideAcc acc[2];
    fr::acc_clear(acc[0]);
    fr::acc_clear(acc[1]);
    uint32_t a[8], b0[8], b1[8];
#pragma unroll
    for (int i = 0; i < 8; i++) {
        a[i] = seed * (2 * i + 1) + threadIdx.x;
        b0[i] = seed * (2 * i + 3) ^ threadIdx.x;
        b1[i] = seed * (2 * i + 5) - threadIdx.x;
    }
    for (int it = 0; it < iters; it++) {
        fr::mul_acc(acc[0], a, b0);
        fr::mul_acc(acc[1], a, b1);
#pragma unroll
        for (int i = 0; i < 8; i++) a[i] += 0x9e3779b9u;  // keep operands changing (ALU pipe, like the LDS refresh)
    }
    uint64_t s = 0;
#pragma unroll
    for (int q = 0; q < 2; q++) {
#pragma unroll
        for (int i = 0; i < 8; i++) s ^= acc[q].e[i];
#pragma unroll
        for (int i = 0; i < 7; i++) s ^= acc[q].o[i];
#pragma unroll
        for (int i = 0; i < 5; i++) s ^= acc[q].ce[i];
#pragma unroll
        for (int i = 0; i < 4; i++) s ^= acc[q].co[i];
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// ---- HBM micro-benchmarks: what the memory system delivers to the access patterns the witness kernels use ---------------
// kind 0: copy (read + write, 16-byte accesses)      -- the pattern behind MEASURED_PEAKS.json hbm_gbs
// kind 1: write-only, streaming 16-byte stores (st.global.cs)
// kind 2: write-only, 256-byte bulk stores from shared memory (cp.async.bulk.global.shared::cta: the witness stream's path)
// kind 3: read-only (16-byte loads, result folded into one word per thread)
__global__ void __launch_bounds__(256) mb_hbm_copy(const uint4* __restrict__ src, uint4* __restrict__ dst, size_t n16) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += stride) __stcs(dst + i, __ldcs(src + i));
}
__global__ void __launch_bounds__(256) mb_hbm_write(uint4* __restrict__ dst, size_t n16, uint32_t seed) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += stride)
        __stcs(dst + i, make_uint4(seed, (uint32_t)i, seed ^ 0x5bd1e995u, (uint32_t)(i >> 32)));
}
__global__ void __launch_bounds__(128) mb_hbm_write_bulk(uint4* __restrict__ dst, size_t n16, uint32_t seed) {
    // every warp owns 32 rows of 256 bytes in shared memory and ships them as 32 bulk copies per round, double-buffered
    extern __shared__ __align__(128) uint4 mb_stage_raw[];
    uint4 (*stage)[128][16] = reinterpret_cast<uint4 (*)[128][16]>(mb_stage_raw);
    const int lane = threadIdx.x & 31;
    const size_t warps = (size_t)gridDim.x * 4, warp = (size_t)blockIdx.x * 4 + (threadIdx.x >> 5);
    const size_t rounds = n16 / (16 * 32);   // 8 KiB per warp-round
    int buf = 0;
    for (size_t r = warp; r < rounds; r += warps, buf ^= 1) {
#pragma unroll
        for (int q = 0; q < 16; q++) stage[buf][threadIdx.x][q] = make_uint4(seed, (uint32_t)r, (uint32_t)q, (uint32_t)lane);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) {
            uint32_t src = static_cast<uint32_t>(__cvta_generic_to_shared(&stage[buf][threadIdx.x][0]));
            uint4* d = dst + r * (16 * 32);
            for (int row = 0; row < 32; row++) {
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], 256;" ::"l"(d), "r"(src) : "memory");
                src += 256;
                d += 16;
            }
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        }
        __syncwarp();
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}
__global__ void __launch_bounds__(256) mb_hbm_read(const uint4* __restrict__ src, uint32_t* __restrict__ sink, size_t n16) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    uint32_t acc = 0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += stride) {
        const uint4 v = __ldcs(src + i);
        acc ^= v.x ^ v.y ^ v.z ^ v.w;
    }
    if (acc == 0x12345678u) sink[0] = acc;   // practically never: keeps the loads alive
}

}  // namespace

int launch_microbench_hbm(h2svd_ctx* ctx, int kind, size_t bytes, double* gb_per_s) {
    if (!gb_per_s || kind < 0 || kind > 3 || bytes < (1u << 20)) {
        set_error("microbench_hbm: bad arguments");
        return H2SVD_EINVAL;
    }
    bytes &= ~(size_t)8191;
    const size_t n16 = bytes / 16;
    H2SVD_TRY(ws_reserve(ctx, (kind == 0 ? 2 : 1) * bytes + 256));
    uint4* buf = (uint4*)ctx->ws;
    uint4* buf2 = buf + n16;
    cudaEvent_t e0, e1;
    H2SVD_CUDA(cudaEventCreate(&e0));
    H2SVD_CUDA(cudaEventCreate(&e1));
    float best_ms = 1e30f;
    const int blocks = ctx->sm_count * 8;
    for (int rep = 0; rep < 6; rep++) {  // rep 0 is the warm-up (and first touch)
        H2SVD_CUDA(cudaEventRecord(e0, ctx->stream));
        switch (kind) {
            case 0: mb_hbm_copy<<<blocks, 256, 0, ctx->stream>>>(buf, buf2, n16); break;
            case 1: mb_hbm_write<<<blocks, 256, 0, ctx->stream>>>(buf, n16, 77u + rep); break;
            case 2:
                H2SVD_SET_SMEM(ctx, mb_hbm_write_bulk, 65536);
                mb_hbm_write_bulk<<<ctx->sm_count * 3, 128, 65536, ctx->stream>>>(buf, n16, 77u + rep);
                break;
            default: mb_hbm_read<<<blocks, 256, 0, ctx->stream>>>(buf, (uint32_t*)(buf + n16), n16); break;
        }
        H2SVD_LAUNCH_CHECK(ctx);
        H2SVD_CUDA(cudaEventRecord(e1, ctx->stream));
        H2SVD_CUDA(cudaEventSynchronize(e1));
        float ms = 0;
        H2SVD_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0 && ms < best_ms) best_ms = ms;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    *gb_per_s = (kind == 0 ? 2.0 : 1.0) * (double)bytes / (best_ms * 1e-3) / 1e9;
    return H2SVD_OK;
}

int launch_microbench(h2svd_ctx* ctx, int kind, int iters, double* ops_per_s) {
    if (!ops_per_s || iters < 1 || kind < 0 || kind > 3) {
        set_error("microbench: bad arguments");
        return H2SVD_EINVAL;
    }
    const int blocks = ctx->sm_count * 8;
    const size_t nthreads = (size_t)blocks * MB_THREADS;
    H2SVD_TRY(ws_reserve(ctx, nthreads * sizeof(uint64_t)));
    cudaEvent_t e0, e1;
    H2SVD_CUDA(cudaEventCreate(&e0));
    H2SVD_CUDA(cudaEventCreate(&e1));
    double per_thread_per_iter = 0;
    float best_ms = 1e30f;
    for (int rep = 0; rep < 4; rep++) {  // rep 0 is the warm-up
        H2SVD_CUDA(cudaEventRecord(e0, ctx->stream));
        switch (kind) {
            case 0:
                mb_imad_lo<<<blocks, MB_THREADS, 0, ctx->stream>>>((uint32_t*)ctx->ws, iters, 12345u + rep);
                per_thread_per_iter = 16.0 * MB_ILP;
                break;
            case 1:
                mb_imad_wide<<<blocks, MB_THREADS, 0, ctx->stream>>>((uint64_t*)ctx->ws, iters, 12345u + rep);
                per_thread_per_iter = 16.0 * MB_ILP;
                break;
            case 2:
                mb_chain<<<blocks, MB_THREADS, 0, ctx->stream>>>((uint64_t*)ctx->ws, iters, 12345u + rep);
                per_thread_per_iter = 8.0 * 4 * 4;
                break;
            default:
                mb_mulacc<<<blocks, MB_THREADS, 0, ctx->stream>>>((uint64_t*)ctx->ws, iters, 12345u + rep);
                per_thread_per_iter = 2.0 * 64;
                break;
        }
        H2SVD_LAUNCH_CHECK(ctx);
        H2SVD_CUDA(cudaEventRecord(e1, ctx->stream));
        H2SVD_CUDA(cudaEventSynchronize(e1));
        float ms = 0;
        H2SVD_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0 && ms < best_ms) best_ms = ms;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    *ops_per_s = per_thread_per_iter * (double)iters * (double)nthreads / (best_ms * 1e-3);
    return H2SVD_OK;
}

}  // namespace h2svd
