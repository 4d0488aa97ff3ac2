// K4 for N-D grids -- FP64 FFT along a STRIDED axis of a contiguous array viewed as
// [outer][N][inner] (the transforms over every axis but the last one of an N-D spectral model:
// demos/nls.ipynb:500-508 `fft2`/`ifft2`, SURVEY.md 8a a14 "2-D 4096^2, 3-D 512^3").
//
// A CTA owns a tile of C adjacent columns (C consecutive `inner` indices = C*16 contiguous bytes per
// row) times all N rows, held row-major in shared memory.  Lane -> (column = lane % C, butterfly):
// the C lanes of a group touch one contiguous C*16-byte segment, both in global memory (coalesced,
// full sectors) and in shared memory (conflict free), so no transpose is ever materialised.
// Levels are the in-place radix-8/16 butterflies of fft_fast.cuh.  As for the rows, the inverse
// transform is decimation-in-frequency (natural order in, digit-reversed order out) and the forward
// transform the mirrored decimation-in-time: physical-space data stay digit-reversed ALONG THIS
// AXIS between the two, which a pointwise nonlinearity does not notice.  First level: global ->
// registers -> tile; last level: tile -> registers -> global; 2 smem round trips for 3 levels.
#pragma once
#include "fft_fast.cuh"

namespace rks {
namespace axis {

// radices, outermost (largest stride) level first; unused levels are 1
template <int N> struct APlan;
template <> struct APlan<16>   { static constexpr int R1 = 16, R2 = 1,  R3 = 1;  };
template <> struct APlan<32>   { static constexpr int R1 = 8,  R2 = 4,  R3 = 1;  };
template <> struct APlan<64>   { static constexpr int R1 = 8,  R2 = 8,  R3 = 1;  };
template <> struct APlan<128>  { static constexpr int R1 = 16, R2 = 8,  R3 = 1;  };
template <> struct APlan<256>  { static constexpr int R1 = 16, R2 = 16, R3 = 1;  };
template <> struct APlan<512>  { static constexpr int R1 = 8,  R2 = 8,  R3 = 8;  };
template <> struct APlan<1024> { static constexpr int R1 = 16, R2 = 8,  R3 = 8;  };
template <> struct APlan<2048> { static constexpr int R1 = 16, R2 = 16, R3 = 8;  };
template <> struct APlan<4096> { static constexpr int R1 = 16, R2 = 16, R3 = 16; };

// tuning knobs of the 512-point tile (tools/build_variant.sh -DRKS_AX512_C=... ; defaults = measured best)
#ifndef RKS_AX512_C
#define RKS_AX512_C 8
#endif
#ifndef RKS_AX512_T
#define RKS_AX512_T 256
#endif
#ifndef RKS_AX512_B
#define RKS_AX512_B 3
#endif
// the same for the 4096-point tile (128 KB at C = 2: one CTA per SM, load / levels / store of a tile do not overlap)
#ifndef RKS_AX4096_C
#define RKS_AX4096_C 2
#endif
#ifndef RKS_AX4096_T
#define RKS_AX4096_T 512
#endif
#ifndef RKS_AX4096_B
#define RKS_AX4096_B 1
#endif
// rows of the NEXT tile that cp.async copies into the spare shared memory while the current tile is transformed
// (0 = off): with one 128 KB tile per SM nothing else overlaps the tile's load with its levels and stores
#ifndef RKS_AX4096_STAGE_ROWS
#define RKS_AX4096_STAGE_ROWS 3072
#endif
#ifndef RKS_AX_UNROLL
#define RKS_AX_UNROLL 1
#endif
// tile geometry: 64 KB tiles (128 KB for N = 4096 so that a row segment is still a full 32-byte sector)
template <int N> RKS_HD constexpr int tile_cols() { return N == 512 ? RKS_AX512_C : N < 512 ? 8 : N == 1024 ? 4 : N == 4096 ? RKS_AX4096_C : 2; }
// threads: one first-level butterfly per thread where the tile has that many (N / R1 butterflies x C columns), so no
// thread idles through a level; resident CTAs per SM: as many as the 227 KB of shared memory and 64 K registers
// allow -- short axes (the second kernel of the two-kernel route, 256^3 grids) need several small tiles in flight
template <int N> RKS_HD constexpr int tile_threads() { return N == 4096 ? RKS_AX4096_T : N == 512 ? RKS_AX512_T : N > 512 ? 256 : N == 256 ? 128 : 64; }
template <int N> RKS_HD constexpr int tile_blocks() { return N == 4096 ? RKS_AX4096_B : N >= 1024 ? 2 : N == 512 ? RKS_AX512_B : N == 256 ? 5 : 8; }
template <int N> RKS_HD constexpr int stage_rows() { return N == 4096 ? RKS_AX4096_STAGE_ROWS : 0; }
template <int N> RKS_HD constexpr int tile_smem() { return (N + stage_rows<N>()) * tile_cols<N>() * 16; }
// persistent CTAs per resident slot: one when the next tile is staged (a CTA's first tile cannot be), else a few
template <int N> RKS_HD constexpr int tile_waves() { return stage_rows<N>() > 0 ? 1 : 4; }
template <int N> RKS_HD constexpr int last_radix() {
    return APlan<N>::R3 > 1 ? APlan<N>::R3 : APlan<N>::R2 > 1 ? APlan<N>::R2 : APlan<N>::R1;
}
// row swizzle (fft_fast.cuh swz): shift 4 when the stride-1 level is radix 16
template <int N> RKS_HD constexpr int tile_shift() { return last_radix<N>() == 16 ? 4 : 3; }

struct Col {                 // one thread's column of the tile and of the global array
    const cplx* gin;         // in  + column offset (row 0)
    cplx* gout;              // out + column offset
    long long gstride;       // elements between consecutive rows inside a row block
    long long bstride;       // elements between consecutive row blocks
    int rb_shift;            // rows per block = 1 << rb_shift (31: a single block, plain strided axis)
    int col;                 // column within the tile
    bool ok;                 // column exists (inner need not be a multiple of C)
    // Scattered output (slab-decomposed grids, dist_fft.py): the rows of the transformed axis belong to different
    // ranks, 1 << o_shift consecutive rows each; otab[g] is the address of rank g's block IN RANK g's MEMORY (peer
    // mapping over NVLink), so the last level's stores ARE the exchange.  nullptr: plain output through gout.
    const long long* otab = nullptr;
    int o_shift = 31;
    long long ooff = 0;      // offset inside a destination block: outer index x block stride + column
    const cplx* stg = nullptr;   // staged copy of the tile's first stage_rows<N>() input rows ([swz(row)][C]), or nullptr
    // Row p of the axis.  Blocked rows: the axis is split over G chunks that sit G-major in memory, as an
    // all-to-all delivers them (dist_fft.py) -- row p = chunk p >> rb_shift, offset p & mask.
    RKS_HD long long row(int p) const {
        return (long long)(p >> rb_shift) * bstride + (long long)(p & (int)((1u << rb_shift) - 1u)) * gstride;
    }
    RKS_HD cplx* out_row(int p) const {
        if (otab)
            return reinterpret_cast<cplx*>(otab[p >> o_shift]) + ooff + (long long)(p & (int)((1u << o_shift) - 1u)) * gstride;
        return gout + row(p);
    }
};

// input row p of the tile's column: from the staged copy when the row was prefetched, else from global memory
template <int N>
RKS_HD cplx first_ld(const Col& c, int p) {
    if (!c.ok) return mk(0.0, 0.0);
    if (stage_rows<N>() > 0 && c.stg && p < stage_rows<N>()) return c.stg[fast::swz<tile_shift<N>()>(p) * tile_cols<N>() + c.col];
    return fast::row_ld(c.gin + c.row(p));
}

// one decimation-in-frequency level of the inverse transform: rows p0 + Q s, twiddles on the outputs
template <int N, int R, int Q, bool FIRST, bool LAST>
RKS_HD void dif_level(cplx* tile, const cplx* tw, const Col& c, int bt, int nbt, double scale) {
    constexpr int C = tile_cols<N>(), SH = tile_shift<N>();
    constexpr int UNR = RKS_AX_UNROLL;
#pragma unroll UNR
    for (int u = bt; u < N / R; u += nbt) {
        const int j = u % Q, p0 = (u / Q) * (R * Q) + j;
        cplx a[R];
#pragma unroll
        for (int s = 0; s < R; ++s) {
            if (FIRST) a[s] = first_ld<N>(c, p0 + Q * s);
            else a[s] = tile[fast::swz<SH>(p0 + Q * s) * C + c.col];
        }
        fast::dftR<R, true>(a);
        if (Q > 1) fast::twiddle_scale<R, true>(a, tw, 0, j * (N / (R * Q)), fast::SlotPerm<R>());
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const cplx v = a[fast::perm<R>(r)];
            if (LAST) { if (c.ok) fast::row_st(c.out_row(p0 + Q * r), mk(v.x * scale, v.y * scale)); }
            else tile[fast::swz<SH>(p0 + Q * r) * C + c.col] = v;
        }
    }
}

// one decimation-in-time level of the forward transform: twiddles on the inputs
template <int N, int R, int Q, bool FIRST, bool LAST>
RKS_HD void dit_level(cplx* tile, const cplx* tw, const Col& c, int bt, int nbt) {
    constexpr int C = tile_cols<N>(), SH = tile_shift<N>();
    constexpr int UNR = RKS_AX_UNROLL;
#pragma unroll UNR
    for (int u = bt; u < N / R; u += nbt) {
        const int j = u % Q, p0 = (u / Q) * (R * Q) + j;
        cplx a[R];
#pragma unroll
        for (int s = 0; s < R; ++s) {
            if (FIRST) a[s] = first_ld<N>(c, p0 + Q * s);
            else a[s] = tile[fast::swz<SH>(p0 + Q * s) * C + c.col];
        }
        if (Q > 1) fast::twiddle_scale<R, false>(a, tw, 0, j * (N / (R * Q)), fast::SlotId());
        fast::dftR<R, false>(a);
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const cplx v = a[fast::perm<R>(r)];
            if (LAST) { if (c.ok) fast::row_st(c.out_row(p0 + Q * r), v); }
            else tile[fast::swz<SH>(p0 + Q * r) * C + c.col] = v;
        }
    }
}

// level LEVEL of one tile.  The caller separates the levels: __syncthreads on the device; the serial host
// emulation runs a level for every thread before the next one
template <int N, bool INV, int LEVEL>
RKS_HD void tile_level(cplx* tile, const cplx* tw, const Col& c, int bt, int nbt, double scale) {
    using P = APlan<N>;
    constexpr int NL = P::R3 > 1 ? 3 : P::R2 > 1 ? 2 : 1;
    constexpr int Q1 = N / P::R1, Q2 = Q1 / P::R2;
    if (LEVEL >= NL) return;
    if (INV) {
        if (LEVEL == 0) dif_level<N, P::R1, Q1, true, NL == 1>(tile, tw, c, bt, nbt, scale);
        else if (LEVEL == 1) dif_level<N, P::R2, Q2, false, NL == 2>(tile, tw, c, bt, nbt, scale);
        else dif_level<N, P::R3, 1, false, true>(tile, tw, c, bt, nbt, scale);
    } else {
        // mirrored order: stride-1 level first, the R1 level last
        if (NL == 1) { dit_level<N, P::R1, Q1, true, true>(tile, tw, c, bt, nbt); return; }
        if (NL == 2) {
            if (LEVEL == 0) dit_level<N, P::R2, Q2, true, false>(tile, tw, c, bt, nbt);
            else dit_level<N, P::R1, Q1, false, true>(tile, tw, c, bt, nbt);
            return;
        }
        if (LEVEL == 0) dit_level<N, P::R3, 1, true, false>(tile, tw, c, bt, nbt);
        else if (LEVEL == 1) dit_level<N, P::R2, Q2, false, false>(tile, tw, c, bt, nbt);
        else dit_level<N, P::R1, Q1, false, true>(tile, tw, c, bt, nbt);
    }
}

}  // namespace axis
}  // namespace rks
