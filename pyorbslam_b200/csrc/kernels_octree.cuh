// K3: deterministic GPU equivalent of ORBextractor::DistributeOctTree (ORBextractor.cpp:539-762) with
// ExtractorNode::DivideNode (:481-537).  One CTA per (image, level); all state is integer and lives in
// shared memory (global scratch only when a level has more FAST candidates than the smem key buffers).
//
// Restated semantics, kept exactly:
//   * node list order: children are push_front'ed in n1..n4 order, the parent is erased; so after a
//     round the list is reverse(children in processing order) ++ untouched nodes in their old order;
//   * full pass: every non-leaf node is divided, in list order;
//   * once |L| + 3*nToExpand > N: ordered rounds -- nodes are divided in descending (size, creation order)
//     (the reference sorts (size, node address); creation order == address order under a monotone allocator,
//     the canonical tie rule) and the round stops as soon as |L| >= N;
//   * result: per node in list order the max-response key, first wins ties; key order inside a node is the
//     stable partition order of the input (cell-row-major, in-cell row-major).
// Parallel form: warp-per-node stable 4-way partition (ballot ranks), then block-wide prefix sums give
// every child its slot in the new list; the ordered rounds rank nodes by counting comparisons.
#pragma once
#include "plan.h"

#ifndef OCT_THREADS
#define OCT_THREADS 256     // measured: 256 threads with three CTAs per SM beat 512 x 2 and 128 x 3 (0.23 vs 0.27 / 0.33 ms per 128 pairs)
#endif
#define OCT_WARPS (OCT_THREADS / 32)

struct OctNodes {      // one generation of the node list (structure of arrays in shared memory)
    short4* box;       // ulx, uly, brx, bry
    int* beg;          // first key
    int* cnt;          // number of keys
    int* meta;         // bit0 leaf (bNoMore), bit1 key buffer, bits 8.. = creation seq + 1 (0 = none)
};

__device__ __forceinline__ int warp_incl_scan(int v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += t;
    }
    return v;
}

// exclusive scan of one value per thread over the block; returns the exclusive prefix, *total = block sum.
// tmp: OCT_WARPS + 1 ints of shared memory.  Contains two __syncthreads().
__device__ __forceinline__ int block_excl_scan(int v, int* tmp, int* total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int incl = warp_incl_scan(v, lane);
    __syncthreads();                      // tmp may still be read from a previous call
    if (lane == 31) tmp[warp] = incl;
    __syncthreads();
    // cross-warp combine: lanes 0..OCT_WARPS-1 of every warp scan the OCT_WARPS warp sums with shuffles
    static_assert(OCT_WARPS <= 32, "one lane per warp sum");
    const int ws = lane < OCT_WARPS ? tmp[lane] : 0;
    const int wincl = warp_incl_scan(ws, lane);
    *total = __shfl_sync(0xffffffffu, wincl, OCT_WARPS - 1);
    const int base = __shfl_sync(0xffffffffu, wincl - ws, warp);
    return base + incl - v;
}

// in-place exclusive scan of a[0..n) (shared memory), every thread owning a contiguous chunk.
__device__ __forceinline__ int block_excl_scan_array(int* a, int n, int* tmp) {
    const int per = (n + OCT_THREADS - 1) / OCT_THREADS;
    const int b = min(threadIdx.x * per, n), e = min(b + per, n);
    int s = 0;
    for (int i = b; i < e; ++i) s += a[i];
    int total;
    int base = block_excl_scan(s, tmp, &total);
    for (int i = b; i < e; ++i) { const int t = a[i]; a[i] = base; base += t; }
    __syncthreads();
    return total;
}

// quadrant of a key inside a node split at (mx, my): 0 = n1 (left/top), 1 = n2, 2 = n3, 3 = n4 (ORBextractor.cpp:509-523)
__device__ __forceinline__ int quadrant(u32 key, int mx, int my) {
    const int x = key & 0xfff, y = (key >> 12) & 0xfff;
    return (x < mx ? 0 : 1) | (y < my ? 0 : 2);
}

// stable 4-way partition of one node's keys by a warp; returns the child sizes (uniform across the warp)
__device__ __forceinline__ int4 warp_partition(const u32* __restrict__ src, u32* __restrict__ dst, int beg, int cnt,
                                               int mx, int my, int lane) {
    int c0 = 0, c1 = 0, c2 = 0, c3 = 0;
    const u32 lt = (1u << lane) - 1;
    if (cnt <= 32) {
        const bool act = lane < cnt;
        const u32 key = act ? src[beg + lane] : 0;
        const int q = act ? quadrant(key, mx, my) : -1;
        const u32 m0 = __ballot_sync(0xffffffffu, q == 0), m1 = __ballot_sync(0xffffffffu, q == 1),
                  m2 = __ballot_sync(0xffffffffu, q == 2), m3 = __ballot_sync(0xffffffffu, q == 3);
        c0 = __popc(m0); c1 = __popc(m1); c2 = __popc(m2); c3 = __popc(m3);
        if (act) {
            const int pos = q == 0 ? __popc(m0 & lt) : q == 1 ? c0 + __popc(m1 & lt)
                          : q == 2 ? c0 + c1 + __popc(m2 & lt) : c0 + c1 + c2 + __popc(m3 & lt);
            dst[beg + pos] = key;
        }
        return make_int4(c0, c1, c2, c3);
    }
    for (int b = 0; b < cnt; b += 32) {
        const int i = b + lane;
        const int q = i < cnt ? quadrant(src[beg + i], mx, my) : -1;
        c0 += __popc(__ballot_sync(0xffffffffu, q == 0));
        c1 += __popc(__ballot_sync(0xffffffffu, q == 1));
        c2 += __popc(__ballot_sync(0xffffffffu, q == 2));
        c3 += __popc(__ballot_sync(0xffffffffu, q == 3));
    }
    int o0 = beg, o1 = beg + c0, o2 = o1 + c1, o3 = o2 + c2;
    for (int b = 0; b < cnt; b += 32) {
        const int i = b + lane;
        const bool act = i < cnt;
        const u32 key = act ? src[beg + i] : 0;
        const int q = act ? quadrant(key, mx, my) : -1;
        const u32 m0 = __ballot_sync(0xffffffffu, q == 0), m1 = __ballot_sync(0xffffffffu, q == 1),
                  m2 = __ballot_sync(0xffffffffu, q == 2), m3 = __ballot_sync(0xffffffffu, q == 3);
        if (act) {
            const int pos = q == 0 ? o0 + __popc(m0 & lt) : q == 1 ? o1 + __popc(m1 & lt)
                          : q == 2 ? o2 + __popc(m2 & lt) : o3 + __popc(m3 & lt);
            dst[pos] = key;
        }
        o0 += __popc(m0); o1 += __popc(m1); o2 += __popc(m2); o3 += __popc(m3);
    }
    return make_int4(c0, c1, c2, c3);
}

// bytes of the node arrays for node capacity capN (two node generations, child counts, rank/order/scan arrays, sort keys)
__host__ __device__ inline size_t oct_node_bytes(int capN) { return (size_t)capN * (2 * (8 + 4 + 4 + 4) + 16 + 4 * 4 + 8); }
// shared-memory bytes of the key buffers, cell offsets and scratch
__host__ __device__ inline size_t oct_base_bytes(int capK, int capC) { return (size_t)2 * capK * 4 + (size_t)capC * 4 + 64 * 4; }

__global__ void __launch_bounds__(OCT_THREADS) k_octree(const __grid_constant__ Plan P, const u32* __restrict__ cand,
                                                        const int* __restrict__ cellcnt, u32* __restrict__ scratch,
                                                        u32* __restrict__ lvl_kp, int* __restrict__ lvl_cnt,
                                                        int capN, int capK, int capC, unsigned char* gnodes, size_t gnode_stride,
                                                        const unsigned char* __restrict__ roottab) {
    pdl_enter();
    extern __shared__ __align__(16) unsigned char oct_smem[];
    const int l = blockIdx.y, slot = blockIdx.x;     // slots fastest: the long level-0 CTAs of all images start first, the short top levels fill the tail
    const LevelGeom& G = P.lv[l];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    // ---- carve memory: key buffers, cell offsets and scratch always in shared memory; the node arrays (80 B per node)
    // follow them in shared memory, or live in a global scratch block when the level quota is too large for that ----
    unsigned char* sp = oct_smem;
    u32* smemKeys0 = (u32*)sp; sp += (size_t)capK * 4;
    u32* smemKeys1 = (u32*)sp; sp += (size_t)capK * 4;
    int* cellOfs = (int*)sp; sp += (size_t)capC * 4;
    int* misc = (int*)sp; sp += 64 * 4;      // [0..32] scan tmp, [40] m
    int* tmp = misc;
    int* sh_m = misc + 40;
    if (gnodes) sp = gnodes + ((size_t)slot * P.nlevels + l) * gnode_stride;     // [slot][level] blocks, b200orb.cu Engine::plan
    unsigned long long* skeys = (unsigned long long*)sp; sp += (size_t)capN * 8;
    int4* cc = (int4*)sp; sp += (size_t)capN * 16;
    short4* boxA = (short4*)sp; sp += (size_t)capN * 8;
    short4* boxB = (short4*)sp; sp += (size_t)capN * 8;
    int* begA = (int*)sp; sp += (size_t)capN * 4;  int* begB = (int*)sp; sp += (size_t)capN * 4;
    int* cntA = (int*)sp; sp += (size_t)capN * 4;  int* cntB = (int*)sp; sp += (size_t)capN * 4;
    int* metaA = (int*)sp; sp += (size_t)capN * 4; int* metaB = (int*)sp; sp += (size_t)capN * 4;
    int* rankOf = (int*)sp; sp += (size_t)capN * 4;
    int* ordIdx = (int*)sp; sp += (size_t)capN * 4;
    int* scanA = (int*)sp; sp += (size_t)capN * 4;
    int* scanB = (int*)sp; sp += (size_t)capN * 4;

    int* out_cnt = lvl_cnt + (size_t)slot * P.nlevels + l;
    u32* out_kp = lvl_kp + (size_t)slot * P.kp_total + G.kp_ofs;

    const int nCells = G.nRows * G.nCols;
    if (nCells <= 0 || G.nIni < 1) { if (tid == 0) *out_cnt = 0; return; }

    // ---- gather the level's candidates in reference order: cell-row-major, in-cell row-major ----
    const int* ccnt = cellcnt + (size_t)slot * P.ncells + G.cell_ofs;
    for (int i = tid; i < nCells; i += OCT_THREADS) cellOfs[i] = ccnt[i];
    __syncthreads();
    const int K = block_excl_scan_array(cellOfs, nCells, tmp);
    if (K == 0) { if (tid == 0) *out_cnt = 0; return; }
    u32* keys[2];
    if (K <= capK) { keys[0] = smemKeys0; keys[1] = smemKeys1; }
    else {   // rare: more candidates than the shared-memory key buffers hold -> same algorithm on global scratch
        u32* base = scratch + ((size_t)slot * P.cand_entries + G.cand_ofs) * 2;
        keys[0] = base; keys[1] = base + (size_t)nCells * G.cell_cap;
    }
    const u32* csrc = cand + (size_t)slot * P.cand_entries + G.cand_ofs;
    // eight lanes per cell (a cell holds ~7 candidates on average, its segment up to cell_cap): four cells per warp step, their loads
    // independent; a cell's count is the difference of its scanned offsets, not another global read
    for (int c = warp * 4 + (lane >> 3); c < nCells; c += OCT_WARPS * 4) {
        const int o = cellOfs[c], n = (c + 1 < nCells ? cellOfs[c + 1] : K) - o;
        for (int k = lane & 7; k < n; k += 8) keys[1][o + k] = csrc[(size_t)c * G.cell_cap + k];
    }
    __syncthreads();

    // ---- roots (ORBextractor.cpp:549-584): stable partition by (int)(x / hX) from keys[1] into keys[0] ----
    // The root of a key comes from a host-built table root_of[x] (the same float division, evaluated once per column), and four
    // roots are counted / scattered per round with 16-bit counters packed in two ints: one pass + one pair of block scans per
    // four roots instead of a pass with a float division and a scan per root.
    const int per = (K + OCT_THREADS - 1) / OCT_THREADS;
    const int kb = min(tid * per, K), ke = min(kb + per, K);
    const unsigned char* rt = roottab + G.root_ofs;
    int nNodes = 0, rootBase = 0;
    // 16-bit counters (4 roots per round) while K < 2^16, else plain 32-bit ones (2 roots per round: the global-scratch path of a
    // level with more candidates than that)
    const bool narrow = K < 65536;
    const int step = narrow ? 4 : 2;
    for (int r0 = 0; r0 < G.nIni; r0 += step) {
        int lo = 0, hi = 0;
        for (int i = kb; i < ke; ++i) {
            const int r = (int)rt[keys[1][i] & 0xfff] - r0;
            if ((unsigned)r >= (unsigned)step) continue;
            if (narrow) { if (r < 2) lo += 1 << (16 * r); else hi += 1 << (16 * (r - 2)); }
            else { if (r == 0) ++lo; else ++hi; }
        }
        int totLo, totHi;
        const int exLo = block_excl_scan(lo, tmp, &totLo);
        const int exHi = block_excl_scan(hi, tmp, &totHi);
        int tot[4], pos[4];
        if (narrow) {
            tot[0] = totLo & 0xffff; tot[1] = (int)((unsigned)totLo >> 16); tot[2] = totHi & 0xffff; tot[3] = (int)((unsigned)totHi >> 16);
            pos[0] = exLo & 0xffff; pos[1] = (int)((unsigned)exLo >> 16); pos[2] = exHi & 0xffff; pos[3] = (int)((unsigned)exHi >> 16);
        } else {
            tot[0] = totLo; tot[1] = totHi; tot[2] = tot[3] = 0;
            pos[0] = exLo; pos[1] = exHi; pos[2] = pos[3] = 0;
        }
        pos[0] += rootBase; pos[1] += rootBase + tot[0]; pos[2] += rootBase + tot[0] + tot[1]; pos[3] += rootBase + tot[0] + tot[1] + tot[2];
        for (int i = kb; i < ke; ++i) {
            const u32 key = keys[1][i];
            const int r = (int)rt[key & 0xfff] - r0;
            if ((unsigned)r < (unsigned)step) {
                int d;
                if (r == 0) d = pos[0]++; else if (r == 1) d = pos[1]++; else if (r == 2) d = pos[2]++; else d = pos[3]++;
                keys[0][d] = key;
            }
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int r = r0 + q;
            if (q < step && r < G.nIni && tot[q] > 0) {
                if (tid == 0) {
                    boxA[nNodes] = make_short4((short)(int)(G.hX * (float)r), 0, (short)(int)(G.hX * (float)(r + 1)),
                                               (short)(G.maxBY - ORB_DET_ORIGIN));
                    begA[nNodes] = rootBase; cntA[nNodes] = tot[q]; metaA[nNodes] = (tot[q] == 1) ? 1 : 0;
                }
                ++nNodes;
                rootBase += tot[q];
            }
        }
    }
    __syncthreads();

    OctNodes cur{boxA, begA, cntA, metaA}, nxt{boxB, begB, cntB, metaB};
    const int N = G.quota;
    int n = nNodes;          // |L|
    bool ordered = false;    // inside the inner while loop of ORBextractor.cpp:672-737
    int E = 0;               // nodes carrying a creation seq (== non-leaf nodes) when ordered

    for (int iter = 0; iter < 64; ++iter) {
        // ---- (A) divide every non-leaf node (speculatively in ordered rounds) ----
        for (int i = warp; i < n; i += OCT_WARPS) {
            const int meta = cur.meta[i];
            if (meta & 1) continue;
            const short4 bx = cur.box[i];
            const int mx = bx.x + ((bx.z - bx.x + 1) >> 1), my = bx.y + ((bx.w - bx.y + 1) >> 1);   // ceil(d/2), :483-484
            const int buf = (meta >> 1) & 1;
            const int4 c = warp_partition(keys[buf], keys[buf ^ 1], cur.beg[i], cur.cnt[i], mx, my, lane);
            if (lane == 0) cc[i] = c;
        }
        __syncthreads();

        // ---- (B1) processing rank of every non-leaf node ----
        if (!ordered) {
            for (int i = tid; i < n; i += OCT_THREADS) scanB[i] = (cur.meta[i] & 1) ? 0 : 1;
            __syncthreads();
            E = block_excl_scan_array(scanB, n, tmp);
            for (int i = tid; i < n; i += OCT_THREADS) {
                const bool leaf = cur.meta[i] & 1;
                rankOf[i] = leaf ? -1 : scanB[i];
                if (!leaf) ordIdx[scanB[i]] = i;
            }
        } else {
            for (int i = tid; i < n; i += OCT_THREADS) {
                const int meta = cur.meta[i];
                rankOf[i] = -1;
                if (!(meta & 1)) {
                    const int seq = (meta >> 8) - 1;
                    skeys[seq] = ((unsigned long long)(unsigned)cur.cnt[i] << 32) | (unsigned)seq;
                    scanB[seq] = i;
                }
            }
            __syncthreads();
            for (int s = tid; s < E; s += OCT_THREADS) {   // descending (size, seq): rank = #keys greater
                const unsigned long long mine = skeys[s];
                int r = 0;
                for (int t = 0; t < E; ++t) r += (skeys[t] > mine);
                const int node = scanB[s];
                rankOf[node] = r;
                ordIdx[r] = node;
            }
        }
        __syncthreads();

        // ---- (B2) children / expandable-children prefix sums in processing order ----
        for (int r = tid; r < E; r += OCT_THREADS) {
            const int4 c = cc[ordIdx[r]];
            const int nch = (c.x > 0) + (c.y > 0) + (c.z > 0) + (c.w > 0);
            const int nex = (c.x > 1) + (c.y > 1) + (c.z > 1) + (c.w > 1);
            scanA[r] = (nch << 16) | nex;
        }
        if (tid == 0) *sh_m = E;
        __syncthreads();
        const int totA = block_excl_scan_array(scanA, E, tmp);
        if (ordered) {   // first rank at which |L| reaches N (the break at ORBextractor.cpp:729-730)
            for (int r = tid; r < E; r += OCT_THREADS) {
                const int4 c = cc[ordIdx[r]];
                const int nch = (c.x > 0) + (c.y > 0) + (c.z > 0) + (c.w > 0);
                if (n + (scanA[r] >> 16) + nch - (r + 1) >= N) atomicMin(sh_m, r + 1);
            }
            __syncthreads();
        }
        const int m = *sh_m;
        const int Ctot = (m == E) ? (totA >> 16) : (scanA[m] >> 16);
        const int Etot = (m == E) ? (totA & 0xffff) : (scanA[m] & 0xffff);

        // ---- (B3) slots of the nodes that stay (leaves and unprocessed), in old list order ----
        for (int i = tid; i < n; i += OCT_THREADS) { const int r = rankOf[i]; scanB[i] = (r < 0 || r >= m) ? 1 : 0; }
        __syncthreads();
        const int stay = block_excl_scan_array(scanB, n, tmp);
        const int newN = Ctot + stay;

        // ---- (B4) build the next list ----
        for (int i = tid; i < n; i += OCT_THREADS) {
            const int r = rankOf[i];
            if (r < 0 || r >= m) {
                const int d = Ctot + scanB[i];
                nxt.box[d] = cur.box[i]; nxt.beg[d] = cur.beg[i]; nxt.cnt[d] = cur.cnt[i]; nxt.meta[d] = cur.meta[i];
            } else {
                const short4 bx = cur.box[i];
                const int mx = bx.x + ((bx.z - bx.x + 1) >> 1), my = bx.y + ((bx.w - bx.y + 1) >> 1);
                const int4 c = cc[i];
                const int cn[4] = {c.x, c.y, c.z, c.w};
                const int buf = ((cur.meta[i] >> 1) & 1) ^ 1;
                int g = scanA[r] >> 16, e = scanA[r] & 0xffff, b = cur.beg[i];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    if (cn[q] > 0) {
                        const int d = Ctot - 1 - g;
                        nxt.box[d] = make_short4((q & 1) ? (short)mx : bx.x, (q & 2) ? (short)my : bx.y,
                                                 (q & 1) ? bx.z : (short)mx, (q & 2) ? bx.w : (short)my);
                        nxt.beg[d] = b; nxt.cnt[d] = cn[q];
                        int meta = (buf << 1) | (cn[q] == 1 ? 1 : 0);
                        if (cn[q] > 1) { meta |= (e + 1) << 8; ++e; }
                        nxt.meta[d] = meta;
                        ++g;
                    }
                    b += cn[q];
                }
            }
        }
        __syncthreads();
        { OctNodes t = cur; cur = nxt; nxt = t; }
        const int prev = n;
        n = newN;
        // ---- (B5) termination, ORBextractor.cpp:668-672 / :733-734 ----
        if (n >= N || n == prev) break;
        if (!ordered && n + 3 * Etot > N) ordered = true;
        E = Etot;
    }

    // ---- result: best response per node, first wins (ORBextractor.cpp:741-759); +16 back to level coords (:836-841) ----
    for (int i = tid; i < n; i += OCT_THREADS) {
        const u32* k = keys[(cur.meta[i] >> 1) & 1] + cur.beg[i];
        u32 best = k[0];
        const int c = cur.cnt[i];
        for (int j = 1; j < c; ++j) { const u32 v = k[j]; if ((v >> 24) > (best >> 24)) best = v; }
        const u32 x = (best & 0xfff) + ORB_DET_ORIGIN, y = ((best >> 12) & 0xfff) + ORB_DET_ORIGIN;
        out_kp[i] = x | (y << 12) | (best & 0xff000000u);
    }
    if (tid == 0) *out_cnt = n;
}
