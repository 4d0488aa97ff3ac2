// K4 building blocks -- power-of-two complex FFT passes on a row held in shared memory, FP64.
//
// The inverse transform is decimation-in-frequency (natural order in, digit-reversed order
// out) and the forward transform is the mirrored decimation-in-time (digit-reversed in,
// natural out).  The pointwise nonlinearity between them does not care about element order,
// so no reordering pass is ever executed.  Passes are radix-4 with one
// radix-2 pass when log2(n) is odd.  A pass is a pure function of (element array, thread
// id, thread count): the kernel separates passes by __syncthreads(); tests/host_check runs
// the same functions serially on the CPU.
//
// Twiddles: tw[j] = exp(-2 pi i j / n), j < n, built once per plan with sincospi.
#pragma once
#include "common.cuh"

namespace rks {

// one radix-4 DIF pass of the INVERSE transform (kernel exp(+2 pi i jk/n)) on blocks of length L
RKS_HD void ifft_dif4_pass(cplx* x, int n, int log2L, const cplx* tw, int tid, int nthreads) {
    const int log2q = log2L - 2;
    const int q = 1 << log2q;
    const int L = 1 << log2L;
    const int tws = n >> log2L;                 // twiddle stride: w_L^j = tw[j * n/L]
    for (int b = tid; b < (n >> 2); b += nthreads) {
        const int blk = b >> log2q, j = b & (q - 1);
        const int base = blk * L + j;
        const cplx a0 = x[base], a1 = x[base + q], a2 = x[base + 2 * q], a3 = x[base + 3 * q];
        const cplx t0 = a0 + a2, t1 = a0 - a2, t2 = a1 + a3, t3 = mul_i(a1 - a3);
        const cplx w1 = conj(tw[j * tws]), w2 = conj(tw[2 * j * tws]), w3 = conj(tw[3 * j * tws]);
        x[base] = t0 + t2;
        x[base + q] = (t1 + t3) * w1;
        x[base + 2 * q] = (t0 - t2) * w2;
        x[base + 3 * q] = (t1 - t3) * w3;
    }
}

// radix-2 pass on blocks of length 2 (twiddle 1): last pass of the DIF chain / first of the DIT chain
RKS_HD void fft_radix2_pass(cplx* x, int n, int tid, int nthreads) {
    for (int b = tid; b < (n >> 1); b += nthreads) {
        const cplx a0 = x[2 * b], a1 = x[2 * b + 1];
        x[2 * b] = a0 + a1;
        x[2 * b + 1] = a0 - a1;
    }
}

// one radix-4 DIT pass of the FORWARD transform (kernel exp(-2 pi i jk/n)) on blocks of length L
RKS_HD void fft_dit4_pass(cplx* x, int n, int log2L, const cplx* tw, int tid, int nthreads) {
    const int log2q = log2L - 2;
    const int q = 1 << log2q;
    const int L = 1 << log2L;
    const int tws = n >> log2L;
    for (int b = tid; b < (n >> 2); b += nthreads) {
        const int blk = b >> log2q, j = b & (q - 1);
        const int base = blk * L + j;
        const cplx b0 = x[base];
        const cplx b1 = x[base + q] * tw[j * tws];
        const cplx b2 = x[base + 2 * q] * tw[2 * j * tws];
        const cplx b3 = x[base + 3 * q] * tw[3 * j * tws];
        const cplx t0 = b0 + b2, t1 = b0 - b2, t2 = b1 + b3, t3 = mul_mi(b1 - b3);
        x[base] = t0 + t2;
        x[base + q] = t1 + t3;
        x[base + 2 * q] = t0 - t2;
        x[base + 3 * q] = t1 - t3;
    }
}

// number of barrier-separated passes of either chain
RKS_HD int fft_num_passes(int log2n) { return (log2n >> 1) + (log2n & 1); }

// pass p (0-based) of the inverse DIF chain: L = n, n/4, ..., then radix-2 if log2n is odd
RKS_HD void ifft_dif_pass(cplx* x, int log2n, int p, const cplx* tw, int tid, int nthreads) {
    const int n = 1 << log2n;
    const int n4 = log2n >> 1;
    if (p < n4) ifft_dif4_pass(x, n, log2n - 2 * p, tw, tid, nthreads);
    else fft_radix2_pass(x, n, tid, nthreads);
}

// pass p (0-based) of the forward DIT chain: exact mirror of the DIF chain
RKS_HD void fft_dit_pass(cplx* x, int log2n, int p, const cplx* tw, int tid, int nthreads) {
    const int n = 1 << log2n;
    const int odd = log2n & 1;
    if (odd && p == 0) { fft_radix2_pass(x, n, tid, nthreads); return; }
    const int k = p - odd;                       // k-th radix-4 pass, block length 4^(k+1) * 2^odd
    fft_dit4_pass(x, n, 2 * (k + 1) + odd, tw, tid, nthreads);
}

}  // namespace rks
