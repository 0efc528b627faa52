// zmarch_plan.h — host-only planning of the z-register marching SpMV (spmv.cu: spmv_zmarch_kernel): how the chains of
// positions are cut into work items.  Plain C++ (no CUDA types), so that the logic is testable on a CPU-only box through
// the hook ab_test_zmarch_plan (tests/test_host_logic.py).
//
// Geometry: M = rows between consecutive chain positions (the plane stride of the stencil), m = rows of the matrix,
// TR = G * 1024 rows per position.  Strip s of the first plane = rows [s * TR, min((s + 1) * TR, M)); position (s, z) =
// that strip shifted by z * M, clipped at m.  Every row of the matrix belongs to exactly one position.  `excl` flags
// 1024-row tiles that must stay out of every item (distributed blocks: tiles that read halo columns; rows with a pattern
// the kernel cannot express); with exclusions the strips must coincide with tiles, i.e. M % TR == 0.
#pragma once
#include <algorithm>
#include <cstdint>
#include <vector>

namespace ab {

constexpr int ZM_TILE_ROWS = 1024;

struct ZmCut {
    int G = 0;          // tiles per position (0: the kernel does not apply)
    int seg_len = 0;    // positions per z-segment (round-robin mode)
    int balanced = 0;   // 1: one contiguous share of the (strip, z) order per CTA
};

struct ZmItems {
    std::vector<uint32_t> first;      // first row of each item's first position
    std::vector<int32_t> count;       // chain length of the item (positions)
    std::vector<int32_t> rows;        // rows per position of the item (<= TR; the last strip of a plane is narrower)
    std::vector<int32_t> cta_first;   // balanced mode: CTA b runs items [cta_first[b], cta_first[b+1]); empty otherwise
    std::vector<int32_t> rest;        // tiles no item covers (only with exclusions), ascending
};

// Picks G and the cut.  Costs are in rows a CTA processes one after the other: round-robin segments cost
// rounds * (L + 2 windows per item) (+ a quarter of the total window count: traffic), a balanced partition (W / grid + 3);
// narrower positions pay a measured penalty (512^3: G = 4 / 2 / 1 -> 0.440 / 0.445 / 0.61 ms), a balanced partition of an
// operand that does not fit the L2 pays for the margins its neighbours no longer share in time.
// fits(g): can the kernel stage positions of g tiles (shared memory)?  g_opt / len_opt / bal_opt: overrides (0 / 0 / -1: auto).
template <class Fits>
inline ZmCut zm_choose_cut(long long M, long long m, int grid, bool any_excl, bool fits_l2, Fits fits, int g_opt, int len_opt, int bal_opt) {
    ZmCut best;
    double bestc = 1e300;
    const long long Zc = (m + M - 1) / M;
    for (int g = 4; g >= 1; g >>= 1) {
        const long long TRg = (long long)g * ZM_TILE_ROWS;
        if (!fits(g)) continue;
        if (any_excl && M % TRg) continue;
        if (g_opt > 0 && g != g_opt) continue;
        if (g_opt <= 0 && g == 1 && best.G) continue;              // single tiles only when nothing wider applies
        const double pen = g == 4 ? 1.0 : (g == 2 ? 1.04 : 1.35);
        const long long Spg = (M + TRg - 1) / TRg;                 // strips per plane
        const double rows = (double)M / (double)Spg;               // average rows per position
        for (int nseg = 1; nseg <= Zc && nseg <= 256; ++nseg) {
            const int Ls = (int)((Zc + nseg - 1) / nseg);
            if (nseg > 1 && (Zc + nseg - 2) / (nseg - 1) == Ls) continue;      // same length as the previous candidate
            if (len_opt > 0 && Ls != len_opt && nseg < Zc) continue;
            const long long nit = Spg * ((Zc + Ls - 1) / Ls);
            if (nit > 16000) break;
            const long long rounds = (nit + grid - 1) / grid;
            const double cost = ((double)rounds * (Ls + 2) + 0.25 * (double)nit * (Ls + 2) / grid) * rows * pen / 1.25;
            if (bal_opt != 1 && cost < bestc) { bestc = cost; best.G = g; best.seg_len = Ls; best.balanced = 0; }
        }
        if (bal_opt != 0) {
            const double per = (double)(Spg * Zc) / grid;
            const double cost = (per + 3.0) * rows * pen * (fits_l2 ? 1.0 : 1.12);
            if (bal_opt == 1 ? (cost < bestc || !best.balanced) : cost < 0.95 * bestc) { bestc = cost; best.G = g; best.seg_len = 0; best.balanced = 1; }
        }
    }
    if (best.G && len_opt > 0 && !best.balanced) best.seg_len = len_opt;
    return best;
}

// The work items of a cut.  excl: one flag per 1024-row tile (empty: none).
inline void zm_make_items(long long M, long long m, const ZmCut& cut, int grid, const std::vector<uint8_t>& excl, ZmItems* out) {
    const long long TR = (long long)cut.G * ZM_TILE_ROWS;
    const long long Sp = (M + TR - 1) / TR, Zc = (m + M - 1) / M;
    const long long ntiles = (m + ZM_TILE_ROWS - 1) / ZM_TILE_ROWS;
    bool any_excl = false;
    for (uint8_t e : excl) any_excl = any_excl || e;
    auto pos_row = [&](long long sI, long long z) { return sI * TR + z * M; };
    auto pos_rows = [&](long long sI, long long z) {
        const long long r0 = pos_row(sI, z);
        long long w = std::min<long long>(TR, M - sI * TR);
        if (r0 + w > m) w = m - r0;
        return w;                                                   // <= 0: the position does not exist
    };
    auto pos_ok = [&](long long sI, long long z) {
        const long long r0 = pos_row(sI, z), w = pos_rows(sI, z);
        if (w <= 0) return false;
        if (any_excl)                                               // strips coincide with tiles here (M % TR == 0)
            for (long long tt = r0 / ZM_TILE_ROWS; tt <= (r0 + w - 1) / ZM_TILE_ROWS; ++tt) if (excl[(size_t)tt]) return false;
        return true;
    };
    std::vector<uint8_t> covered(any_excl ? (size_t)ntiles : 0, 0);
    auto emit = [&](long long sI, long long z0, int run) {
        out->first.push_back((uint32_t)pos_row(sI, z0));
        out->count.push_back(run);
        out->rows.push_back((int32_t)std::min<long long>(TR, M - sI * TR));
        if (any_excl)
            for (int q = 0; q < run; ++q) {
                const long long r0 = pos_row(sI, z0 + q), w = pos_rows(sI, z0 + q);
                for (long long tt = r0 / ZM_TILE_ROWS; tt <= (r0 + w - 1) / ZM_TILE_ROWS; ++tt) covered[(size_t)tt] = 1;
            }
    };
    if (!cut.balanced) {
        const long long L = cut.seg_len > 0 ? cut.seg_len : Zc;
        for (long long z0 = 0; z0 < Zc; z0 += L)
            for (long long sI = 0; sI < Sp; ++sI) {
                int run = 0; long long start = 0;
                for (long long z = z0; z < Zc && z < z0 + L; ++z) {
                    const bool ok = pos_ok(sI, z);
                    if (ok) { if (run == 0) start = z; ++run; }
                    if ((!ok || z + 1 == Zc || z + 1 == z0 + L) && run) { emit(sI, start, run); run = 0; }
                }
            }
    } else {
        // runs of consecutive positions in (strip, z) order, cut into `nc` shares of equal length
        struct Run { long long sI, z0; int len; };
        std::vector<Run> runs;
        long long Wtot = 0;
        for (long long sI = 0; sI < Sp; ++sI) {
            int run = 0; long long start = 0;
            for (long long z = 0; z < Zc; ++z) {
                const bool ok = pos_ok(sI, z);
                if (ok) { if (run == 0) start = z; ++run; }
                if ((!ok || z + 1 == Zc) && run) { runs.push_back({sI, start, run}); Wtot += run; run = 0; }
            }
        }
        const long long nc = Wtot < grid ? Wtot : grid;
        size_t ri = 0; int used = 0;
        for (long long b = 0; b < nc; ++b) {
            out->cta_first.push_back((int32_t)out->first.size());
            long long want = (b + 1) * Wtot / nc - b * Wtot / nc;
            while (want > 0 && ri < runs.size()) {
                const int take = (int)std::min<long long>(want, runs[ri].len - used);
                emit(runs[ri].sI, runs[ri].z0 + used, take);
                used += take; want -= take;
                if (used == runs[ri].len) { ++ri; used = 0; }
            }
        }
        if (nc > 0) out->cta_first.push_back((int32_t)out->first.size());
    }
    if (any_excl) for (long long tI = 0; tI < ntiles; ++tI) if (!covered[(size_t)tI]) out->rest.push_back((int32_t)tI);
}

}  // namespace ab
