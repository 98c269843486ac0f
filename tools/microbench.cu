// Pipe micro-benchmarks for B200 (sm_100a): which instruction mix bounds a PairHMM cell update?
//
// Every kernel runs `iters` iterations of an unrolled body of independent chains in every thread,
// with CTAS_PER_SM x 148 CTAs of 256 threads, and reports warp-instructions / cycle / SM using
// the SM cycle counter (clock64) so the figure is independent of DVFS.
//
// Build:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o microbench microbench.cu
// This is a measurement tool (profiles/microbench_*.txt are its outputs); it is not on the product path.

#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <algorithm>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
    fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)

constexpr int NCHAIN = 8;

struct Out { unsigned long long cycles; float sink; };

#define PROLOGUE \
    float a[NCHAIN], b = seed * 0.999f, c = seed * 1e-3f; \
    _Pragma("unroll") for (int i = 0; i < NCHAIN; ++i) a[i] = seed + i + threadIdx.x; \
    __syncthreads(); \
    unsigned long long t0 = clock64();

#define EPILOGUE \
    unsigned long long t1 = clock64(); \
    float s = 0; _Pragma("unroll") for (int i = 0; i < NCHAIN; ++i) s += a[i]; \
    if (threadIdx.x == 0) { out[blockIdx.x].cycles = t1 - t0; } \
    if (s == 123.456f) out[blockIdx.x].sink = s;

__global__ void k_ffma(Out* out, int iters, float seed) {
    PROLOGUE
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NCHAIN; ++i)
            asm volatile("fma.rn.ftz.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(b), "f"(c));
    }
    EPILOGUE
}

__global__ void k_fmul(Out* out, int iters, float seed) {
    PROLOGUE
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NCHAIN; ++i)
            asm volatile("mul.rn.ftz.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(b));
    }
    (void)c;
    EPILOGUE
}

// packed: each "chain" is a 64-bit register pair
#define PROLOGUE2 \
    unsigned long long a[NCHAIN], b, c; \
    { float2 fb = make_float2(seed * 0.999f, seed * 0.998f), fc = make_float2(seed * 1e-3f, seed * 2e-3f); \
      b = *reinterpret_cast<unsigned long long*>(&fb); c = *reinterpret_cast<unsigned long long*>(&fc); } \
    _Pragma("unroll") for (int i = 0; i < NCHAIN; ++i) { float2 f = make_float2(seed + i + threadIdx.x, seed - i); \
      a[i] = *reinterpret_cast<unsigned long long*>(&f); } \
    __syncthreads(); \
    unsigned long long t0 = clock64();

#define EPILOGUE2 \
    unsigned long long t1 = clock64(); \
    unsigned long long s = 0; _Pragma("unroll") for (int i = 0; i < NCHAIN; ++i) s ^= a[i]; \
    if (threadIdx.x == 0) { out[blockIdx.x].cycles = t1 - t0; } \
    if (s == 0x123456789ull) out[blockIdx.x].sink = 1.f;

__global__ void k_ffma2(Out* out, int iters, float seed) {
    PROLOGUE2
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NCHAIN; ++i)
            asm volatile("fma.rn.ftz.f32x2 %0, %0, %1, %2;" : "+l"(a[i]) : "l"(b), "l"(c));
    }
    EPILOGUE2
}

__global__ void k_fmul2(Out* out, int iters, float seed) {
    PROLOGUE2
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NCHAIN; ++i)
            asm volatile("mul.rn.ftz.f32x2 %0, %0, %1;" : "+l"(a[i]) : "l"(b));
    }
    (void)c;
    EPILOGUE2
}

__global__ void k_fadd2(Out* out, int iters, float seed) {
    PROLOGUE2
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NCHAIN; ++i)
            asm volatile("add.rn.ftz.f32x2 %0, %0, %1;" : "+l"(a[i]) : "l"(c));
    }
    (void)b;
    EPILOGUE2
}

// NALU integer-pipe instructions (LOP3 with predicate out + SEL) interleaved with 8 FFMA2
template <int NALU>
__global__ void k_ffma2_alu(Out* out, int iters, float seed) {
    PROLOGUE2
    unsigned m[NCHAIN];
#pragma unroll
    for (int i = 0; i < NCHAIN; ++i) m[i] = threadIdx.x * 2654435761u + i;
    unsigned hm = (unsigned)seed * 0x11111111u;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NCHAIN; ++i) {
            asm volatile("fma.rn.ftz.f32x2 %0, %0, %1, %2;" : "+l"(a[i]) : "l"(b), "l"(c));
            if (i < NALU)
                asm volatile("{ .reg .pred p; .reg .b32 t; and.b32 t, %0, %1; setp.ne.u32 p, t, 0; selp.b32 %0, %2, %0, p; }"
                             : "+r"(m[i]) : "r"(hm), "r"(it));
        }
    }
    unsigned ms = 0;
#pragma unroll
    for (int i = 0; i < NCHAIN; ++i) ms ^= m[i];
    if (ms == 0x1234567u) out[blockIdx.x].sink = 2.f;
    EPILOGUE2
}

// same with plain FFMA
template <int NALU>
__global__ void k_ffma_alu(Out* out, int iters, float seed) {
    PROLOGUE
    unsigned m[NCHAIN];
#pragma unroll
    for (int i = 0; i < NCHAIN; ++i) m[i] = threadIdx.x * 2654435761u + i;
    unsigned hm = (unsigned)seed * 0x11111111u;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NCHAIN; ++i) {
            asm volatile("fma.rn.ftz.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(b), "f"(c));
            if (i < NALU)
                asm volatile("{ .reg .pred p; .reg .b32 t; and.b32 t, %0, %1; setp.ne.u32 p, t, 0; selp.b32 %0, %2, %0, p; }"
                             : "+r"(m[i]) : "r"(hm), "r"(it));
        }
    }
    unsigned ms = 0;
#pragma unroll
    for (int i = 0; i < NCHAIN; ++i) ms ^= m[i];
    if (ms == 0x1234567u) out[blockIdx.x].sink = 2.f;
    EPILOGUE
}

// FFMA2 with three DISTINCT per-thread register-pair operands (register-bank pressure test)
__global__ void k_ffma2_3r(Out* out, int iters, float seed) {
    PROLOGUE2
    unsigned long long bb[NCHAIN], cc[NCHAIN];
#pragma unroll
    for (int i = 0; i < NCHAIN; ++i) {
        float2 f = make_float2(0.999f - 1e-3f * (threadIdx.x & 7) - 1e-4f * i, 0.998f - 1e-4f * i);
        float2 g = make_float2(1e-3f * i + threadIdx.x * 1e-5f, 2e-3f * i);
        bb[i] = *reinterpret_cast<unsigned long long*>(&f); cc[i] = *reinterpret_cast<unsigned long long*>(&g);
    }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NCHAIN; ++i)
            asm volatile("fma.rn.ftz.f32x2 %0, %0, %1, %2;" : "+l"(a[i]) : "l"(bb[i]), "l"(cc[i]));
    }
    (void)b; (void)c;
    EPILOGUE2
}
// FFMA2 d = a * b + c with a, c distinct per-thread pairs and b distinct too, d != a (4 distinct pairs)
__global__ void k_ffma2_4r(Out* out, int iters, float seed) {
    PROLOGUE2
    unsigned long long bb[NCHAIN], cc[NCHAIN], dd[NCHAIN];
#pragma unroll
    for (int i = 0; i < NCHAIN; ++i) {
        float2 f = make_float2(0.999f - 1e-3f * (threadIdx.x & 7) - 1e-4f * i, 0.998f - 1e-4f * i);
        float2 g = make_float2(1e-3f * i + threadIdx.x * 1e-5f, 2e-3f * i);
        bb[i] = *reinterpret_cast<unsigned long long*>(&f); cc[i] = *reinterpret_cast<unsigned long long*>(&g); dd[i] = 0;
    }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NCHAIN; ++i) {
            asm volatile("fma.rn.ftz.f32x2 %0, %1, %2, %3;" : "=l"(dd[i]) : "l"(a[i]), "l"(bb[i]), "l"(cc[i]));
            asm volatile("mul.rn.ftz.f32x2 %0, %1, %2;" : "=l"(a[i]) : "l"(dd[i]), "l"(bb[(i + 1) % NCHAIN]));
        }
    }
    (void)b; (void)c;
    EPILOGUE2
}
// FMUL2 with two distinct per-thread operands
__global__ void k_fmul2_2r(Out* out, int iters, float seed) {
    PROLOGUE2
    unsigned long long bb[NCHAIN];
#pragma unroll
    for (int i = 0; i < NCHAIN; ++i) {
        float2 f = make_float2(0.999f - 1e-3f * (threadIdx.x & 7) - 1e-4f * i, 0.998f - 1e-4f * i);
        bb[i] = *reinterpret_cast<unsigned long long*>(&f);
    }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NCHAIN; ++i)
            asm volatile("mul.rn.ftz.f32x2 %0, %0, %1;" : "+l"(a[i]) : "l"(bb[i]));
    }
    (void)b; (void)c;
    EPILOGUE2
}
// scalar FFMA with three distinct per-thread operands
__global__ void k_ffma_3r(Out* out, int iters, float seed) {
    PROLOGUE
    float bb[NCHAIN], cc[NCHAIN];
#pragma unroll
    for (int i = 0; i < NCHAIN; ++i) { bb[i] = 0.999f - 1e-3f * (threadIdx.x & 7) - 1e-4f * i; cc[i] = 1e-3f * i + threadIdx.x * 1e-5f; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NCHAIN; ++i)
            asm volatile("fma.rn.ftz.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(bb[i]), "f"(cc[i]));
    }
    (void)b; (void)c;
    EPILOGUE
}

// shuffles only
__global__ void k_shfl(Out* out, int iters, float seed) {
    PROLOGUE
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NCHAIN; ++i)
            a[i] = __shfl_up_sync(0xffffffffu, a[i], 1);
    }
    (void)b; (void)c;
    EPILOGUE
}

// NSH shuffles per 8 FFMA2
template <int NSH>
__global__ void k_ffma2_shfl(Out* out, int iters, float seed) {
    PROLOGUE2
    float sh[NCHAIN];
#pragma unroll
    for (int i = 0; i < NCHAIN; ++i) sh[i] = seed + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NCHAIN; ++i) {
            asm volatile("fma.rn.ftz.f32x2 %0, %0, %1, %2;" : "+l"(a[i]) : "l"(b), "l"(c));
            if (i < NSH) sh[i] = __shfl_up_sync(0xffffffffu, sh[i], 1);
        }
    }
    float ss = 0;
#pragma unroll
    for (int i = 0; i < NCHAIN; ++i) ss += sh[i];
    if (ss == 1.2345f) out[blockIdx.x].sink = ss;
    EPILOGUE2
}

__global__ void k_dfma(Out* out, int iters, float seed) {
    double a[NCHAIN], b = seed * 0.999, c = seed * 1e-3;
#pragma unroll
    for (int i = 0; i < NCHAIN; ++i) a[i] = seed + i + threadIdx.x;
    __syncthreads();
    unsigned long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NCHAIN; ++i)
            asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(a[i]) : "d"(b), "d"(c));
    }
    unsigned long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < NCHAIN; ++i) s += a[i];
    if (threadIdx.x == 0) out[blockIdx.x].cycles = t1 - t0;
    if (s == 123.456) out[blockIdx.x].sink = (float)s;
}

// shared-memory word loads interleaved with FFMA2 (1 LDS per 8 FFMA2)
__global__ void k_ffma2_lds(Out* out, int iters, float seed) {
    __shared__ unsigned sm[2048];
    for (int i = threadIdx.x; i < 2048; i += blockDim.x) sm[i] = i * 7u;
    PROLOGUE2
    unsigned acc = 0, idx = threadIdx.x;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NCHAIN; ++i)
            asm volatile("fma.rn.ftz.f32x2 %0, %0, %1, %2;" : "+l"(a[i]) : "l"(b), "l"(c));
        acc ^= sm[(idx + it) & 2047];
    }
    if (acc == 0x7654321u) out[blockIdx.x].sink = 3.f;
    EPILOGUE2
}

template <typename K>
static void run(const char* name, K kernel, int warp_instr_per_iter, int threads, int ctas_per_sm, int sms, int iters) {
    int grid = sms * ctas_per_sm;
    Out* d; CK(cudaMalloc(&d, grid * sizeof(Out)));
    CK(cudaMemset(d, 0, grid * sizeof(Out)));
    kernel<<<grid, threads>>>(d, iters / 8, 1.0f);   // warm-up
    CK(cudaDeviceSynchronize());
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    CK(cudaEventRecord(e0));
    kernel<<<grid, threads>>>(d, iters, 1.0f);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    std::vector<Out> h(grid);
    CK(cudaMemcpy(h.data(), d, grid * sizeof(Out), cudaMemcpyDeviceToHost));
    std::vector<unsigned long long> cyc(grid);
    for (int i = 0; i < grid; ++i) cyc[i] = h[i].cycles;
    std::sort(cyc.begin(), cyc.end());
    double med = (double)cyc[grid / 2];
    double warps_per_sm = (double)threads / 32 * ctas_per_sm;
    double winstr_per_sm = warps_per_sm * (double)iters * warp_instr_per_iter;
    // all CTAs of one SM run concurrently, so SM time ~= CTA time
    double ipc = winstr_per_sm / med;
    double ghz = med / (ms * 1e6);
    printf("%-28s thr=%4d cta/sm=%d  warp-instr/clk/SM = %6.3f  (lane-ops/clk/SM = %7.1f)  med_cycles=%.0f  ms=%.3f  ~%.3f GHz\n",
           name, threads, ctas_per_sm, ipc, ipc * 32, med, ms, ghz);
    CK(cudaFree(d));
}

int main(int argc, char** argv) {
    int dev = 0; CK(cudaSetDevice(dev));
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, dev));
    int sms = p.multiProcessorCount;
    printf("device: %s  SMs=%d  cc=%d.%d  clockRate=%d kHz\n", p.name, sms, p.major, p.minor, p.clockRate);
    int iters = 20000;
    for (int cfg = 0; cfg < 3; ++cfg) {
        int threads = cfg == 0 ? 128 : (cfg == 1 ? 256 : 512);
        int cps = cfg == 0 ? 1 : 2;      // 4, 16, 32 warps per SM
        printf("--- %d warps/SM ---\n", threads / 32 * cps);
        run("ffma (3-reg)", k_ffma, NCHAIN, threads, cps, sms, iters);
        run("fmul", k_fmul, NCHAIN, threads, cps, sms, iters);
        run("ffma2 (f32x2)", k_ffma2, NCHAIN, threads, cps, sms, iters);
        run("ffma2 3 distinct reg pairs", k_ffma2_3r, NCHAIN, threads, cps, sms, iters);
        run("ffma2(4 distinct)+fmul2", k_ffma2_4r, 2 * NCHAIN, threads, cps, sms, iters);
        run("fmul2 2 distinct reg pairs", k_fmul2_2r, NCHAIN, threads, cps, sms, iters);
        run("ffma 3 distinct regs", k_ffma_3r, NCHAIN, threads, cps, sms, iters);
        run("fmul2", k_fmul2, NCHAIN, threads, cps, sms, iters);
        run("fadd2", k_fadd2, NCHAIN, threads, cps, sms, iters);
        run("8 ffma2 + 2x(lop,setp,sel)", k_ffma2_alu<2>, NCHAIN, threads, cps, sms, iters);
        run("8 ffma2 + 4x(lop,setp,sel)", k_ffma2_alu<4>, NCHAIN, threads, cps, sms, iters);
        run("8 ffma2 + 8x(lop,setp,sel)", k_ffma2_alu<8>, NCHAIN, threads, cps, sms, iters);
        run("8 ffma + 2x(lop,setp,sel)", k_ffma_alu<2>, NCHAIN, threads, cps, sms, iters);
        run("8 ffma + 4x(lop,setp,sel)", k_ffma_alu<4>, NCHAIN, threads, cps, sms, iters);
        run("8 ffma + 8x(lop,setp,sel)", k_ffma_alu<8>, NCHAIN, threads, cps, sms, iters);
        run("shfl.up only", k_shfl, NCHAIN, threads, cps, sms, iters);
        run("8 ffma2 + 1 shfl", k_ffma2_shfl<1>, NCHAIN, threads, cps, sms, iters);
        run("8 ffma2 + 2 shfl", k_ffma2_shfl<2>, NCHAIN, threads, cps, sms, iters);
        run("8 ffma2 + 4 shfl", k_ffma2_shfl<4>, NCHAIN, threads, cps, sms, iters);
        run("8 ffma2 + 1 lds", k_ffma2_lds, NCHAIN, threads, cps, sms, iters);
        run("dfma", k_dfma, NCHAIN, threads, cps, sms, iters / 4);
    }
    printf("note: 'warp-instr' counts only the FMA-class instructions of the body (8 per iteration);\n"
           "      lane-ops for f32x2 kernels are packed instructions (x2 FMAs each).\n");
    return 0;
}
