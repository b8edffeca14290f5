// Pipe / cache peak microbenchmarks: the roofline denominators MEASURED_PEAKS.json does not carry.
//
// The driver measures HBM copy bandwidth and cuBLAS bf16 on this pool; the kernels of this path are bound by other
// units: fp32 FMA issue (kNN D=3, loss), the ALU pipe (top-k selection), fp64 FMA / the
// fp64 tensor cores (decode's cosine matrix), the TF32 tensor pipe (kNN D=63 Gram) and L2 row gathers (loss, re-rank).
// bench.py launches each of these for a few milliseconds on the same GPU, in the same process, and divides the
// kernels' achieved rates by what it measured here.  Every kernel is a dependency-free register / L2 loop sized to keep
// the unit saturated; `work` (flops or bytes per launch) is computed on the host from the same constants.
#include "common.cuh"
#include "tc_ptx.cuh"

namespace hpcs {

constexpr int kPeakThreads = 256;
constexpr int kPeakCtasPerSm = 8;

// ---- 0: packed fp32 FMA (FFMA2), 8 independent float2 chains per thread --------------------------------------
__global__ void __launch_bounds__(kPeakThreads) peak_ffma2_kernel(int iters, float* out) {
    unsigned long long acc[8];
    const float s = 1.0f + 1e-7f * threadIdx.x;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const float2 v = make_float2(s + i, s - i);
        acc[i] = *reinterpret_cast<const unsigned long long*>(&v);
    }
    const float2 m2 = make_float2(0.9999f, 1.0001f), a2 = make_float2(1e-3f, -1e-3f);
    const unsigned long long m = *reinterpret_cast<const unsigned long long*>(&m2), a = *reinterpret_cast<const unsigned long long*>(&a2);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int i = 0; i < 8; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(acc[i]) : "l"(m), "l"(a));
    }
    float r = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const float2 v = *reinterpret_cast<const float2*>(&acc[i]);
        r += v.x + v.y;
    }
    if (r == 123.456f) out[0] = r;
}
// ---- 1: scalar fp32 FMA (three-register form), 16 chains ---------------------------------------------------------
__global__ void __launch_bounds__(kPeakThreads) peak_ffma_kernel(int iters, float* out) {
    float acc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = 1.0f + i + 1e-7f * threadIdx.x;
    float m = 0.9999f + 1e-9f * threadIdx.x, a = 1e-3f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 2; ++u)
#pragma unroll
            for (int i = 0; i < 16; ++i) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(acc[i]) : "f"(m), "f"(a));
    }
    float r = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) r += acc[i];
    if (r == 123.456f) out[0] = r;
}
// ---- 2: ALU pipe (fp32 min/max, what the top-k selection networks issue), 16 chains --------------------------------
__global__ void __launch_bounds__(kPeakThreads) peak_alu_kernel(int iters, float* out) {
    float acc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = 1.0f + i + 1e-3f * threadIdx.x;
    float lo = 0.5f + threadIdx.x, hi = 7.5f + threadIdx.x;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            asm volatile("max.f32 %0, %0, %1;" : "+f"(acc[i]) : "f"(lo));
            asm volatile("min.f32 %0, %0, %1;" : "+f"(acc[i]) : "f"(hi));
        }
    }
    float r = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) r += acc[i];
    if (r == 123.456f) out[0] = r;
}
// ---- 3: fp64 FMA, 8 chains ----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kPeakThreads) peak_dfma_kernel(int iters, float* out) {
    double acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 1.0 + i + 1e-9 * threadIdx.x;
    double m = 0.9999 + 1e-12 * threadIdx.x, a = 1e-3;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 2; ++u)
#pragma unroll
            for (int i = 0; i < 8; ++i) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(acc[i]) : "d"(m), "d"(a));
    }
    double r = 0.;
#pragma unroll
    for (int i = 0; i < 8; ++i) r += acc[i];
    if (r == 123.456) out[0] = (float)r;
}
// ---- 4: fp64 tensor cores (mma.sync m8n8k4, what pdist_mma issues), 4 accumulator fragments -----------------------------------
__global__ void __launch_bounds__(kPeakThreads) peak_dmma_kernel(int iters, float* out) {
    double c[4][2];
#pragma unroll
    for (int i = 0; i < 4; ++i) c[i][0] = c[i][1] = 1e-3 * i;
    double a = 1.0 + 1e-9 * threadIdx.x, b = 0.999;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int i = 0; i < 4; ++i)
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                             : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
    double r = 0.;
#pragma unroll
    for (int i = 0; i < 4; ++i) r += c[i][0] + c[i][1];
    if (r == 123.456) out[0] = (float)r;
}
// ---- 5: TF32 tcgen05.mma, M=128 N=256 K=8, operands resident in shared memory, accumulator in TMEM -----------------------------
constexpr int kUmmaN = 256;
constexpr int kUmmaBatch = 64;                                   // MMAs between commits
__global__ void __launch_bounds__(128) peak_umma_tf32_kernel(int iters, float* out) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    float* a_tile = reinterpret_cast<float*>(base);                               // 128 rows x 128 B
    float* b_tile = reinterpret_cast<float*>(base + 128 * 128);                   // 256 rows x 128 B
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    for (int i = threadIdx.x; i < (128 + kUmmaN) * 32; i += blockDim.x) a_tile[i] = 1.0f + (float)((i * 2654435761u) >> 20) * 1e-4f;
    if (threadIdx.x == 0) { ptx::mbar_init(&bar, 1); ptx::fence_barrier_init(); }
    if (threadIdx.x < 32) ptx::tmem_alloc<kUmmaN>(&tmem_slot);
    ptx::fence_proxy_async_smem();
    ptx::tc_fence_before_sync();
    __syncthreads();
    ptx::tc_fence_after_sync();
    const uint32_t tmem = tmem_slot;
    if (threadIdx.x == 0) {
        const uint32_t idesc = ptx::umma_idesc_tf32(128, kUmmaN);
        const uint32_t a_addr = ptx::smem_u32(a_tile), b_addr = ptx::smem_u32(b_tile);
        uint32_t parity = 0;
        for (int it = 0; it < iters; ++it) {
#pragma unroll 4
            for (int j = 0; j < kUmmaBatch; ++j) {
                const int kk = j & 3;
                ptx::umma_tf32(tmem, ptx::umma_desc_k_sw128(a_addr + kk * 32), ptx::umma_desc_k_sw128(b_addr + kk * 32), idesc,
                               (it | j) != 0);
            }
            ptx::umma_commit(&bar);
            ptx::mbar_wait(&bar, parity);
            parity ^= 1;
        }
    }
    ptx::tc_fence_before_sync();
    __syncthreads();
    ptx::tc_fence_after_sync();
    if (threadIdx.x < 32) {
        float v[32];
        ptx::tmem_ld_32x32(tmem, v);
        if (v[0] == 123.456f) out[0] = v[1];
    }
    ptx::tc_fence_before_sync();
    __syncthreads();
    if (threadIdx.x < 32) ptx::tmem_dealloc<kUmmaN>(tmem);
}
// ---- 6/7: L2 row gathers: random 128-byte (8 lanes) or 256-byte (16 lanes) rows of an L2-resident table ---------------------------
template <int LANES>
__global__ void __launch_bounds__(kPeakThreads) peak_l2_gather_kernel(const float4* __restrict__ table, unsigned rows, int iters, float* out) {
    const unsigned gid = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned group = gid / LANES, sub = gid % LANES;
    unsigned state = group * 2654435761u + 12345u;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int it = 0; it < iters; ++it) {
        float4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {                                  // 4 independent rows in flight per group
            state = state * 1664525u + 1013904223u;
            const unsigned row = __umulhi(state, rows);
            v[u] = __ldg(table + (size_t)row * LANES + sub);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) { acc.x += v[u].x; acc.y += v[u].y; acc.z += v[u].z; acc.w += v[u].w; }
    }
    if (acc.x + acc.y + acc.z + acc.w == 123.456f) out[0] = acc.x;
}

}  // namespace hpcs

// which: 0 ffma2, 1 ffma, 2 alu (min/max), 3 dfma, 4 dmma, 5 umma tf32, 6 L2 gather 128 B rows, 7 L2 gather 256 B rows.
// work_host (optional, HOST pointer) receives the flops (0-5: 2 per FMA; 2: one per min/max) or bytes (6, 7) of this launch.
extern "C" int hpcs_peak_probe(int which, int iters, const void* table, size_t table_bytes, float* out, double* work_host,
                               void* stream) {
    using namespace hpcs;
    if (iters <= 0 || !out) return fail(HPCS_ERR_ARG, "peak_probe: bad arguments");
    const int sms = sm_count();
    const int grid = sms * kPeakCtasPerSm;
    const double threads = (double)grid * kPeakThreads;
    cudaStream_t st = as_stream(stream);
    double work = 0;
    switch (which) {
        case 0: peak_ffma2_kernel<<<grid, kPeakThreads, 0, st>>>(iters, out); work = threads * iters * 4.0 * 8 * 4; break;
        case 1: peak_ffma_kernel<<<grid, kPeakThreads, 0, st>>>(iters, out); work = threads * iters * 2.0 * 16 * 2; break;
        case 2: peak_alu_kernel<<<grid, kPeakThreads, 0, st>>>(iters, out); work = threads * iters * 16.0 * 2; break;
        case 3: peak_dfma_kernel<<<grid, kPeakThreads, 0, st>>>(iters, out); work = threads * iters * 2.0 * 8 * 2; break;
        case 4: peak_dmma_kernel<<<grid, kPeakThreads, 0, st>>>(iters, out); work = threads / 32 * iters * 16.0 * 512; break;
        case 5: {
            const int smem = (128 + kUmmaN) * 128 + 1024;
            static thread_local bool attr_set = false;
            if (!attr_set) {
                cudaFuncSetAttribute(peak_umma_tf32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
                attr_set = true;
            }
            peak_umma_tf32_kernel<<<sms, 128, smem, st>>>(iters, out);
            work = (double)sms * iters * kUmmaBatch * 2.0 * 128 * kUmmaN * 8;
            break;
        }
        case 6:
        case 7: {
            const int lanes = which == 6 ? 8 : 16;
            const size_t rows = table_bytes / (lanes * 16);
            if (!table || rows < 2 || rows > 0xffffffffu) return fail(HPCS_ERR_ARG, "peak_probe: table too small");
            if (which == 6) peak_l2_gather_kernel<8><<<grid, kPeakThreads, 0, st>>>((const float4*)table, (unsigned)rows, iters, out);
            else peak_l2_gather_kernel<16><<<grid, kPeakThreads, 0, st>>>((const float4*)table, (unsigned)rows, iters, out);
            work = threads * iters * 4.0 * 16;
            break;
        }
        default: return fail(HPCS_ERR_ARG, "peak_probe: unknown probe %d", which);
    }
    if (work_host) *work_host = work;
    return check_launch("peak_probe");
}
