// membench.cu -- what the memory system alone allows for the CartPole step's access pattern:
// per env read 4 x f64 state + i32 counter + i64 action (44 B), write 4 x f64 + i32 + float4 obs +
// f64 reward + 2 x u8 (62 B), no arithmetic to speak of.  Build: nvcc -arch=sm_100a -O3 membench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <vector>
#include <algorithm>

struct Args { double* s; int* el; long long* act; float4* obs; double* rew; uint8_t* te; uint8_t* tr; int64_t n; };

template <int MODE>  // 0 = read+write, 1 = read only, 2 = write only
__global__ void __launch_bounds__(256) pattern(Args a) {
    int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= a.n) return;
    double s0 = 1, s1 = 2, s2 = 3, s3 = 4; int el = 0; long long ac = 0;
    if (MODE != 2) {
        s0 = a.s[i]; s1 = a.s[a.n + i]; s2 = a.s[2 * a.n + i]; s3 = a.s[3 * a.n + i];
        el = a.el[i]; ac = a.act[i];
    }
    s0 += 1.0; s1 += s0; s2 += s1; s3 += s2 + (double)ac; el += 1;
    if (MODE != 1) {
        a.s[i] = s0; a.s[a.n + i] = s1; a.s[2 * a.n + i] = s2; a.s[3 * a.n + i] = s3; a.el[i] = el;
        a.obs[i] = make_float4((float)s0, (float)s1, (float)s2, (float)s3);
        a.rew[i] = 1.0; a.te[i] = el & 1; a.tr[i] = (el >> 1) & 1;
    } else if (s3 == 12345.678) a.rew[i] = s3;
}

__global__ void fill(unsigned char* p, size_t n, unsigned char v) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; size_t st = (size_t)gridDim.x * blockDim.x;
    for (; i < n / 16; i += st) ((uint4*)p)[i] = make_uint4(v, v, v, v);
}
__global__ void rd(const uint4* p, size_t n, unsigned* out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; size_t st = (size_t)gridDim.x * blockDim.x;
    unsigned acc = 0; for (; i < n; i += st) { uint4 v = p[i]; acc += v.x ^ v.y ^ v.z ^ v.w; }
    if (acc == 0xdeadbeef) *out = acc;
}

template <int MODE> float run(Args a, int reps, bool flush, unsigned char* fb, unsigned char* fb2, unsigned* sink) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    std::vector<float> t;
    for (int r = 0; r < reps; r++) {
        if (flush) { fill<<<1184, 256>>>(fb, 256u << 20, (unsigned char)r); rd<<<1184, 256>>>((uint4*)fb2, (256u << 20) / 16, sink); }
        cudaEventRecord(e0);
        pattern<MODE><<<(unsigned)((a.n + 255) / 256), 256>>>(a);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); t.push_back(ms);
    }
    std::sort(t.begin(), t.end());
    return t[t.size() / 2];
}

int main() {
    unsigned char *fb, *fb2; unsigned* sink;
    cudaMalloc(&fb, 256u << 20); cudaMalloc(&fb2, 256u << 20); cudaMalloc(&sink, 4); cudaMemset(fb2, 0, 256u << 20);
    for (int lg : {20, 22, 24}) {
        Args a; a.n = 1ll << lg;
        cudaMalloc(&a.s, 32 * a.n); cudaMalloc(&a.el, 4 * a.n); cudaMalloc(&a.act, 8 * a.n); cudaMalloc(&a.obs, 16 * a.n);
        cudaMalloc(&a.rew, 8 * a.n); cudaMalloc(&a.te, a.n); cudaMalloc(&a.tr, a.n);
        cudaMemset(a.s, 0, 32 * a.n); cudaMemset(a.el, 0, 4 * a.n); cudaMemset(a.act, 0, 8 * a.n);
        for (int flush = 0; flush < 2; flush++) {
            float rw = run<0>(a, 40, flush, fb, fb2, sink), r = run<1>(a, 40, flush, fb, fb2, sink), w = run<2>(a, 40, flush, fb, fb2, sink);
            printf("n=2^%d flush=%d  read+write %.2f us (%.0f GB/s of 106 B/env)   read-only %.2f us (%.0f GB/s of 44)   write-only %.2f us (%.0f GB/s of 62)\n",
                   lg, flush, rw * 1e3, 106.0 * a.n / rw / 1e6, r * 1e3, 44.0 * a.n / r / 1e6, w * 1e3, 62.0 * a.n / w / 1e6);
        }
        cudaFree(a.s); cudaFree(a.el); cudaFree(a.act); cudaFree(a.obs); cudaFree(a.rew); cudaFree(a.te); cudaFree(a.tr);
    }
    return 0;
}
