// ILP probe for the Cash-Karp attempt (tuning experiment, not part of the product): attempts per second with one
// instance per thread vs two instances advanced in lockstep by the same thread, at the occupancy each one allows.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o gpurun_out/ilp_probe tools/ilp_probe.cu
#include <cstdio>
#include "../cosmotransitions_b200/csrc/ctb_device.cuh"
using namespace ctb;

template <int NI, int MINB>
__global__ void __launch_bounds__(128, MINB) probe(const double *c, double *out, int iters)
{
    Sfi<PotPoly4, 2> s[NI];
    double y0[NI], y1[NI], d0[NI], d1[NI], t[NI], dt[NI];
#pragma unroll
    for (int i = 0; i < NI; i++) {
        const int id = (blockIdx.x * blockDim.x + threadIdx.x) * NI + i;
        const double cc = c[id % 4096];
        const double p[5] = {0, 0, cc / 2, -(1 + cc) / 3, 0.25};
        s[i].pot.load(p, nullptr, nullptr);
        s[i].phi_abs = 1.0; s[i].phi_meta = 0.0; s[i].n_rkck = 0;
        y0[i] = 0.9 - 1e-4 * i; y1[i] = -0.01; t[i] = 1.0 + 0.01 * (id & 7); dt[i] = 0.05;
        s[i].eom(y0[i], y1[i], t[i], d0[i], d1[i]);
    }
    for (int k = 0; k < iters; k++) { // one ATTEMPT per trip for every instance of the thread (lockstep state machine)
        double dy0[NI], dy1[NI], e0[NI], e1[NI];
#pragma unroll
        for (int i = 0; i < NI; i++) s[i].rkck(y0[i], y1[i], d0[i], d1[i], t[i], dt[i], dy0[i], dy1[i], e0[i], e1[i]);
#pragma unroll
        for (int i = 0; i < NI; i++) {
            // error norm and controller of rkqs (HF:113-179), branch-free form
            const double a0 = fabs(e0[i]) * 1e4, b0 = fabs(e0[i]) * rcp_fast((fabs(y0[i]) + 1e-300) * 1e-4);
            const double a1 = fabs(e1[i]) * 1e4, b1 = fabs(e1[i]) * rcp_fast((fabs(y1[i]) + 1e-300) * 1e-4);
            const double errmax = fmax(fmin(a0, b0), fmin(a1, b1));
            const bool ok = errmax < 1.0;
            const double grow = (errmax > 1.89e-4) ? 0.9 * inv_fifth_root(fmin(errmax, 1.0)) : 5.0;
            const double shrink = fmax(0.1, 0.9 * rsqrt(sqrt(fmax(errmax, 1.0))));
            const double r1 = t[i] + dt[i], na = y0[i] + dy0[i], nb = y1[i] + dy1[i];
            double nda, ndb;
            s[i].eom_inv(na, nb, s[i].inv_r_end, nda, ndb);
            const bool wrap = na < 0.5 || na > 1.0 || r1 > 50.0;
            if (ok) {
                y0[i] = wrap ? 0.9 : na; y1[i] = wrap ? -0.01 : nb; t[i] = wrap ? 1.0 : r1;
                d0[i] = nda; d1[i] = ndb;
            }
            dt[i] = fmin(fmax(dt[i] * (ok ? grow : shrink), 1e-3), 0.2);
        }
    }
    double acc = 0;
#pragma unroll
    for (int i = 0; i < NI; i++) acc += y0[i] + y1[i] + s[i].n_rkck;
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int NI, int MINB>
void run(const char *name, int blocks_per_sm)
{
    const int nblk = 148 * blocks_per_sm, iters = 20000;
    double *c, *out;
    cudaMalloc(&c, 4096 * 8); cudaMalloc(&out, nblk * 128 * 8);
    double h[4096];
    for (int i = 0; i < 4096; i++) h[i] = 0.02 + 0.475 * (i + 0.5) / 4096;
    cudaMemcpy(c, h, sizeof(h), cudaMemcpyHostToDevice);
    probe<NI, MINB><<<nblk, 128>>>(c, out, 100);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    probe<NI, MINB><<<nblk, 128>>>(c, out, iters);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    cudaFuncAttributes a; cudaFuncGetAttributes(&a, probe<NI, MINB>);
    printf("%-28s regs %3d  blocks/SM %d  %.3f ms  %.3e attempts/s (err %s)\n", name, a.numRegs, blocks_per_sm, ms,
           (double)nblk * 128 * NI * iters / (ms * 1e-3), cudaGetErrorString(cudaGetLastError()));
    cudaFree(c); cudaFree(out);
}

int main()
{
    run<1, 3>("1 instance/thread, 3 blk/SM", 3);
    run<1, 4>("1 instance/thread, 4 blk/SM", 4);
    run<1, 6>("1 instance/thread, 6 blk/SM", 6);
    run<2, 2>("2 instances/thread, 2 blk/SM", 2);
    run<2, 3>("2 instances/thread, 3 blk/SM", 3);
    run<2, 4>("2 instances/thread, 4 blk/SM", 4);
    run<4, 1>("4 instances/thread, 1 blk/SM", 1);
    run<4, 2>("4 instances/thread, 2 blk/SM", 2);
    run<4, 3>("4 instances/thread, 3 blk/SM", 3);
    return 0;
}
