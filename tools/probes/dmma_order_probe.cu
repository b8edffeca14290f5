// Which order does mma.sync.m8n8k4.f64 add its four products in?  (decides whether the fp64 tensor cores can
// reproduce scipy's running sums bit for bit).  Operands are fp32 values widened to fp64, like the decoder's.
//   nvcc -gencode arch=compute_100a,code=sm_100a -o /tmp/dmma_probe tools/probes/dmma_order_probe.cu && /tmp/dmma_probe
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cuda_runtime.h>

__global__ void probe(const double* A, const double* B, const double* C, double* D) {
    const int lane = threadIdx.x;
    const double a = A[(lane / 4) * 4 + lane % 4];            // A[row = lane/4][k = lane%4]
    const double b = B[(lane % 4) * 8 + lane / 4];            // B[k = lane%4][col = lane/4]
    const int r = lane / 4, c0 = 2 * (lane % 4);
    double c[2] = {C[r * 8 + c0], C[r * 8 + c0 + 1]}, d[2];
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%4,%5};"
                 : "=d"(d[0]), "=d"(d[1]) : "d"(a), "d"(b), "d"(c[0]), "d"(c[1]));
    D[r * 8 + c0] = d[0];
    D[r * 8 + c0 + 1] = d[1];
}

int main() {
    double hA[32], hB[32], hC[64], hD[64];
    int match[4] = {0, 0, 0, 0}, total = 0;
    double *dA, *dB, *dC, *dD;
    cudaMalloc(&dA, sizeof hA); cudaMalloc(&dB, sizeof hB); cudaMalloc(&dC, sizeof hC); cudaMalloc(&dD, sizeof hD);
    srand(1);
    for (int trial = 0; trial < 2000; ++trial) {
        auto rnd = [&]() { return (double)(float)((rand() / (double)RAND_MAX - 0.5) * (trial % 3 == 0 ? 1e-3 : 2.0)); };
        for (double& v : hA) v = rnd();
        for (double& v : hB) v = rnd();
        for (double& v : hC) v = trial % 2 ? rnd() * rnd() * 7.0 : 0.0;
        cudaMemcpy(dA, hA, sizeof hA, cudaMemcpyHostToDevice); cudaMemcpy(dB, hB, sizeof hB, cudaMemcpyHostToDevice);
        cudaMemcpy(dC, hC, sizeof hC, cudaMemcpyHostToDevice);
        probe<<<1, 32>>>(dA, dB, dC, dD);
        if (cudaMemcpy(hD, dD, sizeof hD, cudaMemcpyDeviceToHost) != cudaSuccess) { printf("cuda error\n"); return 1; }
        for (int i = 0; i < 8; ++i) for (int j = 0; j < 8; ++j) {
            const double* a = hA + i * 4;
            double bk[4] = {hB[0 * 8 + j], hB[1 * 8 + j], hB[2 * 8 + j], hB[3 * 8 + j]};
            double asc = hC[i * 8 + j], desc = hC[i * 8 + j];
            for (int k = 0; k < 4; ++k) asc = fma(a[k], bk[k], asc);
            for (int k = 3; k >= 0; --k) desc = fma(a[k], bk[k], desc);
            const double tree = (a[0] * bk[0] + a[1] * bk[1]) + (a[2] * bk[2] + a[3] * bk[3]) + hC[i * 8 + j];   // products exact
            long double ex = (long double)hC[i * 8 + j];
            for (int k = 0; k < 4; ++k) ex += (long double)a[k] * (long double)bk[k];
            const double got = hD[i * 8 + j];
            match[0] += got == asc; match[1] += got == desc; match[2] += got == tree; match[3] += got == (double)ex;
            ++total;
        }
    }
    printf("outputs %d: == fma chain k ascending %d, k descending %d, pairwise tree %d, single rounding of the exact sum %d\n",
           total, match[0], match[1], match[2], match[3]);
    return 0;
}
