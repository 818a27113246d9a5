// Developer micro-benchmark (not part of the library): what the memory system delivers to the WRITE PATTERN of the
// witness stream -- 32 lanes of a warp each own a stripe of W x 32 bytes (consecutive stripes are consecutive in memory)
// and fill it CH witnesses (CH x 32 bytes) at a time -- for several store mechanisms, burst sizes and residencies, with
// no arithmetic at all.  Measured on B200 (W = 60, 1 M elements = 2.0 GB per launch; profiles/r02_store_pattern.txt):
//   * 32 stripes x 256-byte pieces per burst (the rescale kernel's pattern):       5.3 TB/s at 4 to 12 warps per SM
//   * the same with W = 56 or 64 (every piece 256-byte ALIGNED):                   5.9-6.1 TB/s
//   * W = 60 with every lane's pieces cut at the 256-byte boundaries of memory:    5.9-6.2 TB/s (mode 5, per-row bulk copies),
//                                                                                   5.6-5.9 TB/s (mode 6, one TMA box per chunk)
//   * a warp's burst as one contiguous piece (not the required layout, mode 3):    6.2-6.3 TB/s = the write peak
//   * direct 16-byte stores from registers into the 32 stripes (mode 1):           1.5-2.2 TB/s
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o variants/store_pattern tools/store_pattern.cu
//   mode 0: per-row bulk copies from shared memory (cp.async.bulk), double-buffered   (the rescale kernel's path)
//   mode 1: direct 16-byte stores from registers, each lane into its own stripe (no staging)
//   mode 2: staged, then coalesced 16-byte stores (half a warp per 256-byte piece)
//   mode 3: like mode 0 but every burst of a warp is ONE contiguous piece of memory (not the required layout: shows what
//           the scatter over 32 stripes costs)
//   mode 4: like mode 0, single-buffered
//   mode 6: TMA tensor stores with 256-byte-aligned pieces (what the library ships, csrc/rescale_dev.cuh WitnessStreamTma): the
//           even and the odd stripes of a warp are two 3-D tensors [128 B][W*32/128 halves][elements / 2]; each class of 16
//           lanes ships its completed 8-witness chunk as ONE box {128 B, 2 halves, 16 rows}; the odd class starts with a
//           one-half head box (a store with a negative coordinate faults); needs W % 8 == 4
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x)                                                                                  \
    do {                                                                                       \
        cudaError_t e_ = (x);                                                                  \
        if (e_ != cudaSuccess) {                                                               \
            fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_));         \
            exit(1);                                                                           \
        }                                                                                      \
    } while (0)

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool elect_one() {
    uint32_t is_leader;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(is_leader));
    return is_leader != 0;
}

template <int MODE>
__global__ void __launch_bounds__(256) store_kernel(uint4* __restrict__ out, size_t elements, int W, int CH, int nbuf) {
    extern __shared__ __align__(16) uint4 stage[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int row_u4 = 2 * CH + 1;
    const uint32_t buf_stride = blockDim.x * row_u4;
    uint4* row0 = stage + (size_t)threadIdx.x * row_u4;
    uint4* warp_row0 = stage + (size_t)(threadIdx.x - lane) * row_u4;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    int buf = 0;
    for (size_t e0 = (size_t)blockIdx.x * blockDim.x + warp * 32; e0 < elements; e0 += stride) {  // elements is a multiple of 32
        uint4* gwarp = out + e0 * (size_t)W * 2;   // lane 0's stripe, in 16-byte units
        for (int w0 = 0; w0 < W; w0 += CH) {
            const int fill = W - w0 < CH ? W - w0 : CH;
            const uint4 v = make_uint4((uint32_t)e0 + lane, (uint32_t)w0, 0x9e3779b9u, (uint32_t)lane);
            if (MODE == 1) {
                uint4* d = gwarp + (size_t)lane * W * 2 + (size_t)w0 * 2;
                for (int i = 0; i < 2 * fill; i++) __stcs(d + i, v);
                continue;
            }
            uint4* s = row0 + buf * buf_stride;
            for (int i = 0; i < 2 * fill; i++) s[i] = v;
            if (MODE == 2) {
                __syncwarp();
                const int chunk = lane & 15, half = lane >> 4;
                if (chunk < 2 * fill)
                    for (int r = half; r < 32; r += 2)
                        __stcs(gwarp + (size_t)r * W * 2 + (size_t)w0 * 2 + chunk, warp_row0[r * row_u4 + chunk]);
                __syncwarp();
                continue;
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (elect_one()) {
                const uint32_t bytes = (uint32_t)fill * 32u;
                uint32_t src = smem_addr(warp_row0 + buf * buf_stride);
                uint4* dst = MODE == 3 ? gwarp + (size_t)w0 * 2 * 32 : gwarp + (size_t)w0 * 2;
                const size_t dstep = MODE == 3 ? (size_t)fill * 2 : (size_t)W * 2;
                for (int r = 0; r < 32; r++) {
                    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes)
                                 : "memory");
                    src += row_u4 * 16;
                    dst += dstep;
                }
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                if (nbuf == 2)
                    asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                else
                    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            }
            __syncwarp();
            if (nbuf == 2) buf ^= 1;
        }
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// mode 5: the required layout, 8-witness bursts, but every lane's chunk boundaries sit on 256-BYTE-ALIGNED addresses of
// its stripe (stripes of W x 32 bytes start 0, 32, .. 224 bytes past such a boundary).  Per lane a ring of two 8-witness
// buffers; witness w of an element goes to ring position (8 * F0 + a + w) & 15 (a = the stripe's misalignment in
// witnesses, F0 = the warp's flush counter at the start of the element), so that aligned chunks coincide with buffers and
// flush number F always ships buffer F & 1.  A flush happens every 8 witnesses and at the end of the element; each lane
// publishes what it ships (destination, source, bytes) in a 16-byte descriptor and one elected lane issues the copies.
__global__ void __launch_bounds__(128) store_aligned_kernel(uint4* __restrict__ out, size_t elements, int W) {
    extern __shared__ __align__(16) uint4 stage[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int ROW_U4 = 33;
    uint4* row = stage + (size_t)threadIdx.x * ROW_U4;
    const uint32_t row_s = smem_addr(row);
    uint4* desc = stage + (size_t)blockDim.x * ROW_U4 + warp * 32;     // this warp's descriptor table
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    uint32_t F = 0;            // flushes so far (warp-uniform)
    int wait_at = 8;           // puts after a flush before which the previous flush's reads must be complete
    for (size_t e0 = (size_t)blockIdx.x * blockDim.x + warp * 32; e0 < elements; e0 += stride) {
        uint4* mine = out + (e0 + lane) * (size_t)W * 2;
        const uint32_t a = (uint32_t)((reinterpret_cast<uintptr_t>(mine) >> 5) & 7u);
        uint32_t amax = a;
        for (int o = 16; o; o >>= 1) amax = max(amax, __shfl_xor_sync(0xffffffffu, amax, o));
        const uint32_t F0 = F;
        int fill = 0;
        uint32_t j = 0;   // chunks of this element shipped so far
        auto ship = [&](uint32_t chunk, uint32_t count) {   // chunk index within the element; `count` witnesses exist
            const uint32_t lo = max(8u * chunk, a), hi = min(min(8u * chunk + 8u, (uint32_t)W + a), count + a);
            const uint32_t bytes = hi > lo ? (hi - lo) * 32u : 0u;
            uint4* dst = mine + (size_t)(lo - a) * 2;
            const uint32_t src = row_s + (((8u * F0 + lo) & 15u) * 32u);
            desc[lane] = make_uint4((uint32_t)reinterpret_cast<uintptr_t>(dst), (uint32_t)(reinterpret_cast<uintptr_t>(dst) >> 32), src, bytes);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (elect_one()) {
                for (int r = 0; r < 32; r++) {
                    const uint4 d = desc[r];
                    if (d.w != 0u) {
                        const unsigned long long g = ((unsigned long long)d.y << 32) | d.x;
                        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(g), "r"(d.z), "r"(d.w) : "memory");
                    }
                }
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                if (amax == 0u) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
            }
            __syncwarp();
            F++;
        };
        wait_at = amax == 0u ? 8 : 8 - (int)amax;
        for (int w = 0; w < W; w++) {
            const uint4 v = make_uint4((uint32_t)(e0 + lane), (uint32_t)w, 0x9e3779b9u, (uint32_t)lane);
            if (fill == wait_at && amax != 0u) {
                asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                __syncwarp();
            }
            uint4* s = row + 2 * ((8u * F0 + a + (uint32_t)w) & 15u);
            s[0] = v;
            s[1] = v;
            if (++fill == 8) {
                ship(j++, (uint32_t)w + 1u);
                fill = 0;
            }
        }
        // end of the element: the remaining chunks (one, or two when the last witnesses straddle a boundary for some lane).
        // If the last flush period was too short to reach its wait point, the flush before it has not been waited for yet.
        const uint32_t nchunks = ((uint32_t)W + amax + 7u) / 8u;
        const bool two = nchunks - j >= 2u;
        if (amax != 0u && fill <= wait_at) {
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            __syncwarp();
        }
        while (j < nchunks) ship(j++, (uint32_t)W);
        if (two) {   // both buffers in flight: the next element's first witnesses would overwrite the older one
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            __syncwarp();
        }
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

typedef CUresult (*tma_encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// the library's WitnessStreamTma (csrc/rescale_dev.cuh) without the arithmetic
__global__ void __launch_bounds__(128) store_tma_kernel(const __grid_constant__ CUtensorMap map_even, const __grid_constant__ CUtensorMap map_odd,
                                                        const __grid_constant__ CUtensorMap map_head, size_t elements, int W) {
    extern __shared__ uint8_t raw[];
    const uint32_t raw_s = smem_addr(raw);
    uint8_t* stage = raw + (((raw_s + 1023u) & ~1023u) - raw_s);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // per warp: [class 2][buffer 2] tiles of 16 rows x 256 B (4 KB each) + the 16 x 128 B head tile of the odd class
    uint8_t* wbase = stage + (size_t)warp * 18432;
    const uint32_t cls = lane & 1, rho = lane >> 1;
    uint32_t J0 = 0, J1 = 0;                    // chunks each class has shipped so far (buffer parity), warp-uniform
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t e0 = (size_t)blockIdx.x * blockDim.x + warp * 32; e0 < elements; e0 += stride) {
        const int row0 = (int)(e0 >> 1);
        uint32_t done0 = 0, done1 = 0;
        for (uint32_t count = 0; count < (uint32_t)W;) {
            const uint4 v = make_uint4((uint32_t)(e0 + lane), count, 0x9e3779b9u, (uint32_t)lane);
            uint8_t* line;
            uint32_t sw, c;
            if (cls == 1u && count < 4u) {
                line = wbase + 4 * 4096 + rho * 128;
                sw = rho & 7u;
                c = 2u * count;
            } else {
                const uint32_t pos = cls ? (8u * J1 + count - 4u) & 15u : (8u * J0 + count) & 15u;
                const uint32_t buf = pos >> 3, s = pos & 7u, half = s >> 2;
                line = wbase + (cls * 2 + buf) * 4096 + rho * 256 + half * 128;
                sw = (2u * rho + half) & 7u;
                c = 2u * (s & 3u);
            }
            *reinterpret_cast<uint4*>(line + ((c ^ sw) << 4)) = v;
            *reinterpret_cast<uint4*>(line + (((c + 1u) ^ sw) << 4)) = v;
            count++;
            const uint32_t ph = count & 7u;
            if (ph == 0u || ph == 4u) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (elect_one()) {
                    if (ph == 0u) {
                        const uint32_t src = smem_addr(wbase + (0 * 2 + ((J0 + done0) & 1u)) * 4096);
                        asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%1, %2, %3}], [%4];" ::"l"(&map_even),
                                     "r"(0), "r"(2 * (int)done0), "r"(row0), "r"(src)
                                     : "memory");
                    } else if (count == 4u) {
                        const uint32_t src = smem_addr(wbase + 4 * 4096);
                        asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%1, %2, %3}], [%4];" ::"l"(&map_head),
                                     "r"(0), "r"(0), "r"(row0), "r"(src)
                                     : "memory");
                    } else {
                        const uint32_t src = smem_addr(wbase + (1 * 2 + ((J1 + done1) & 1u)) * 4096);
                        asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%1, %2, %3}], [%4];" ::"l"(&map_odd),
                                     "r"(0), "r"(2 * (int)done1 + 1), "r"(row0), "r"(src)
                                     : "memory");
                    }
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    asm volatile("cp.async.bulk.wait_group.read 2;" ::: "memory");
                }
                __syncwarp();
                if (ph == 0u) done0++;
                else if (count != 4u) done1++;
            }
        }
        // end of the stripe (W % 8 == 4): the even class has 4 witnesses left (the second half of its box is past the stripe: clipped)
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (elect_one()) {
            const uint32_t src = smem_addr(wbase + (0 * 2 + ((J0 + done0) & 1u)) * 4096);
            asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%1, %2, %3}], [%4];" ::"l"(&map_even), "r"(0),
                         "r"(2 * (int)done0), "r"(row0), "r"(src)
                         : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 2;" ::: "memory");
        }
        __syncwarp();
        J0 += done0 + 1u;
        J1 += done1;
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

int main(int argc, char** argv) {
    const int W = argc > 1 ? atoi(argv[1]) : 60;
    const size_t elements = 1024 * 1024;
    const size_t bytes = elements * W * 32;
    uint4* out;
    CK(cudaMalloc(&out, bytes));
    uint8_t* flush;
    CK(cudaMalloc(&flush, 256u << 20));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    struct Cfg { int mode, ch, nbuf, threads, ctas; };
    std::vector<Cfg> cfgs;
    for (int ch : {4, 8, 12, 16, 20})
        for (int warps : {4, 8, 12}) cfgs.push_back({0, ch, 2, 128, warps / 4});
    for (int warps : {8, 12}) cfgs.push_back({3, 8, 2, 128, warps / 4});
    for (int warps : {8, 16, 32}) cfgs.push_back({1, 8, 1, 128, warps / 4});
    for (int ctas : {1, 2, 3}) {
        const size_t smem = 128 * 33 * 16 + 4 * 32 * 16;
        CK(cudaFuncSetAttribute(store_aligned_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        float best = 1e30f;
        for (int rep = 0; rep < 4; rep++) {
            CK(cudaMemsetAsync(flush, rep, 256u << 20));
            CK(cudaEventRecord(e0));
            store_aligned_kernel<<<sms * ctas, 128, smem>>>(out, elements, W);
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            CK(cudaGetLastError());
            float ms;
            CK(cudaEventElapsedTime(&ms, e0, e1));
            if (rep > 0 && ms < best) best = ms;
        }
        printf("W %d mode 5 (aligned chunks) CH  8 ring 16 128 thr x %d ctas/SM: %7.1f us  %6.0f GB/s\n", W, ctas, best * 1e3,
               bytes / (best * 1e-3) / 1e9);
        // verify the pattern landed where it should (lane/witness ids)
        if (ctas == 3) {
            std::vector<uint4> h(elements * W * 2);
            CK(cudaMemcpy(h.data(), out, bytes, cudaMemcpyDeviceToHost));
            size_t bad = 0;
            for (size_t e = 0; e < elements; e++)
                for (int w = 0; w < W; w++)
                    for (int q = 0; q < 2; q++) {
                        const uint4 v = h[(e * W + w) * 2 + q];
                        if (v.x != (uint32_t)e || v.y != (uint32_t)w || v.w != (uint32_t)(e & 31)) bad++;
                    }
            printf("W %d mode 5 check: %zu bad of %zu\n", W, bad, elements * W * 2);
        }
        fflush(stdout);
    }
    if (W % 8 == 4) {
        void* fnp = nullptr;
        cudaDriverEntryPointQueryResult qres;
        CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &qres));
        tma_encode_fn encode = reinterpret_cast<tma_encode_fn>(fnp);
        CUtensorMap maps[3];   // even stripes, odd stripes, the first 128 bytes of the odd stripes
        for (int c = 0; c < 3; c++) {
            const int odd = c >= 1;
            const cuuint64_t dims[3] = {128, (cuuint64_t)(W * 32 / 128), (cuuint64_t)(elements / 2)};
            const cuuint64_t strides[2] = {128, (cuuint64_t)W * 64};
            const cuuint32_t box[3] = {128, c == 2 ? 1u : 2u, 16};
            const cuuint32_t estr[3] = {1, 1, 1};
            const CUresult r = encode(&maps[c], CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, (uint8_t*)out + (size_t)odd * W * 32, dims, strides, box, estr,
                                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) {
                fprintf(stderr, "tensor map %d failed: %d\n", c, (int)r);
                return 1;
            }
        }
        for (int ctas : {1, 2, 3}) {
            const size_t smem = 4 * 18432 + 1024;
            CK(cudaFuncSetAttribute(store_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            float best = 1e30f;
            for (int rep = 0; rep < 4; rep++) {
                CK(cudaMemsetAsync(flush, rep, 256u << 20));
                if (rep == 1 && ctas == 3) CK(cudaMemsetAsync(out, 0xff, bytes));
                CK(cudaEventRecord(e0));
                store_tma_kernel<<<sms * ctas, 128, smem>>>(maps[0], maps[1], maps[2], elements, W);
                CK(cudaEventRecord(e1));
                CK(cudaEventSynchronize(e1));
                CK(cudaGetLastError());
                float ms;
                CK(cudaEventElapsedTime(&ms, e0, e1));
                if (rep > 0 && ms < best) best = ms;
            }
            printf("W %d mode 6 (TMA boxes, aligned, two classes) 128 thr x %d ctas/SM: %7.1f us  %6.0f GB/s\n", W, ctas, best * 1e3,
                   bytes / (best * 1e-3) / 1e9);
            if (ctas == 3) {
                std::vector<uint4> h(elements * W * 2);
                CK(cudaMemcpy(h.data(), out, bytes, cudaMemcpyDeviceToHost));
                size_t bad = 0;
                for (size_t e = 0; e < elements; e++)
                    for (int w = 0; w < W; w++)
                        for (int q = 0; q < 2; q++) {
                            const uint4 v = h[(e * W + w) * 2 + q];
                            if (v.x != (uint32_t)e || v.y != (uint32_t)w || v.w != (uint32_t)(e & 31)) bad++;
                        }
                printf("W %d mode 6 check: %zu bad of %zu\n", W, bad, elements * W * 2);
            }
            fflush(stdout);
        }
    }
    for (const Cfg& c : cfgs) {
        const int nbuf = c.nbuf;
        const size_t smem = c.mode == 1 ? 0 : (size_t)c.threads * nbuf * (2 * c.ch + 1) * 16;
        if ((smem + 1024) * c.ctas > 227 * 1024) continue;
        void (*kern)(uint4*, size_t, int, int, int) = c.mode == 0 ? store_kernel<0> : c.mode == 1 ? store_kernel<1> : store_kernel<3>;
        CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        float best = 1e30f;
        for (int rep = 0; rep < 4; rep++) {
            CK(cudaMemsetAsync(flush, rep, 256u << 20));
            CK(cudaEventRecord(e0));
            kern<<<sms * c.ctas, c.threads, smem>>>(out, elements, W, c.ch, nbuf);
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            CK(cudaGetLastError());
            float ms;
            CK(cudaEventElapsedTime(&ms, e0, e1));
            if (rep > 0 && ms < best) best = ms;
        }
        printf("W %d mode %d CH %2d nbuf %d %3d thr x %2d ctas/SM (%2d warps, %3zu KB staging): %7.1f us  %6.0f GB/s\n", W, c.mode, c.ch, nbuf,
               c.threads, c.ctas, c.ctas * c.threads / 32, smem * c.ctas / 1024, best * 1e3, bytes / (best * 1e-3) / 1e9);
        fflush(stdout);
    }
    return 0;
}
