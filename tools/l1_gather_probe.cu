// l1_gather_probe.cu -- what does a divergent LDG.128 / LDG.256 cost in the L1 data stage of a B200 SM?
//
// Tool only (not part of libpyvr_cuda.so).  The march kernel (csrc/march.cu) is bound by the L1 data stage: its four
// corner-row loads per sample are warp-wide gathers of 32-byte (f32x4 z-pair) or 16-byte (f16x4 z-pair) entries.  ncu
// shows ~12 data-stage wavefronts per LDG.256 where 8 would move the bytes.  This probe measures the rule directly:
// every lane of a warp loads VEC bytes from  window + 128*line[lane] + VEC*slot[lane]  over and over (all hits after
// the first pass), for address patterns read from stdin, and reports SM cycles per warp-wide load.
//
//   stdin:  one pattern per line:  <name> <vec bytes: 4|8|16|32> <32 x "line:slot">
//   stdout: name, vec, cycles per warp-load per SM (clock64 of the slowest CTA), ns per warp-load per SM (events)
//
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_bin/l1_gather_probe tools/l1_gather_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); exit(1); } } while (0)

constexpr int THREADS = 256, CTAS_PER_SM = 4, WINDOW = 32768, UNROLL = 8;

template <int VEC>
__device__ __forceinline__ unsigned load(const char *p) {
    if constexpr (VEC == 32) {
        unsigned long long a, b, c, d;
        asm volatile("ld.global.nc.v4.b64 {%0, %1, %2, %3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p));
        const unsigned long long x = a ^ b ^ c ^ d; return (unsigned)x ^ (unsigned)(x >> 32);
    } else if constexpr (VEC == 16) {
        unsigned a, b, c, d;
        asm volatile("ld.global.nc.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "l"(p));
        return a ^ b ^ c ^ d;
    } else if constexpr (VEC == 8) {
        unsigned a, b;
        asm volatile("ld.global.nc.v2.b32 {%0, %1}, [%2];" : "=r"(a), "=r"(b) : "l"(p));
        return a ^ b;
    } else {
        unsigned a;
        asm volatile("ld.global.nc.b32 %0, [%1];" : "=r"(a) : "l"(p));
        return a;
    }
}

template <int VEC>
__global__ void __launch_bounds__(THREADS, CTAS_PER_SM)
probe(const char *buf, const int *lane_off, int iters, unsigned stride, unsigned *sink, long long *cycles) {
    const char *w = buf + (size_t)blockIdx.x * WINDOW + lane_off[threadIdx.x & 31];
    unsigned acc = 0;
    // warm the window (every address of the pattern, both halves)
    acc ^= load<VEC>(w) ^ load<VEC>(w + WINDOW / 2);
    __syncthreads();
    const long long t0 = clock64();
    for (int k = 0; k < iters; ++k) {
#pragma unroll
        for (int u = 0; u < UNROLL; ++u)   // `stride` = WINDOW / 2 at run time: the two halves alternate, nothing is loop-invariant
            acc ^= load<VEC>(w + (((unsigned)(k * UNROLL + u) * stride) & (WINDOW - 1)));
    }
    const long long t1 = clock64();
    if (acc == 0x12345678u) *sink = acc;
    if ((threadIdx.x & 31) == 0) atomicMax((unsigned long long *)(cycles + blockIdx.x), (unsigned long long)(t1 - t0));
}

int main(int argc, char **argv) {
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    const int grid = sms * CTAS_PER_SM, iters = argc > 1 ? atoi(argv[1]) : 4096;
    char *buf; int *d_off; unsigned *sink; long long *d_cycles;
    CK(cudaMalloc(&buf, (size_t)grid * WINDOW));
    CK(cudaMemset(buf, 1, (size_t)grid * WINDOW));
    CK(cudaMalloc(&d_off, 32 * sizeof(int)));
    CK(cudaMalloc(&sink, sizeof(unsigned)));
    CK(cudaMalloc(&d_cycles, grid * sizeof(long long)));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    std::vector<long long> h_cycles(grid);
    char line[4096];
    printf("%-44s %4s %10s %10s\n", "pattern", "vec", "cyc/load", "ns/load");
    while (fgets(line, sizeof line, stdin)) {
        if (line[0] == '#' || line[0] == '\n') { if (line[0] == '#') fputs(line, stdout); continue; }
        char name[256]; int vec = 0, pos = 0;
        if (sscanf(line, "%255s %d%n", name, &vec, &pos) != 2) continue;
        int off[32]; const char *p = line + pos; bool ok = true;
        for (int l = 0; l < 32; ++l) {
            int ln, sl, n;
            if (sscanf(p, " %d:%d%n", &ln, &sl, &n) != 2) { ok = false; break; }
            p += n;
            off[l] = ln * 128 + sl * vec;
            if (off[l] + vec > WINDOW / 2) ok = false;
        }
        if (!ok) { fprintf(stderr, "bad pattern: %s", line); continue; }
        CK(cudaMemcpy(d_off, off, sizeof off, cudaMemcpyHostToDevice));
        float best_ms = 1e30f; long long best_cyc = 1LL << 62;
        for (int rep = 0; rep < 3; ++rep) {
            CK(cudaMemset(d_cycles, 0, grid * sizeof(long long)));
            CK(cudaEventRecord(e0));
            switch (vec) {
                case 32: probe<32><<<grid, THREADS>>>(buf, d_off, iters, WINDOW / 2, sink, d_cycles); break;
                case 16: probe<16><<<grid, THREADS>>>(buf, d_off, iters, WINDOW / 2, sink, d_cycles); break;
                case 8: probe<8><<<grid, THREADS>>>(buf, d_off, iters, WINDOW / 2, sink, d_cycles); break;
                default: probe<4><<<grid, THREADS>>>(buf, d_off, iters, WINDOW / 2, sink, d_cycles); break;
            }
            CK(cudaGetLastError());
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
            CK(cudaMemcpy(h_cycles.data(), d_cycles, grid * sizeof(long long), cudaMemcpyDeviceToHost));
            long long mx = 0; for (long long c : h_cycles) if (c > mx) mx = c;
            if (ms < best_ms) best_ms = ms;
            if (mx < best_cyc) best_cyc = mx;
        }
        const double loads_per_sm = (double)CTAS_PER_SM * (THREADS / 32) * iters * UNROLL;
        printf("%-44s %4d %10.2f %10.3f\n", name, vec, best_cyc / loads_per_sm, best_ms * 1e6 / loads_per_sm);
        fflush(stdout);
    }
    return 0;
}
