// hb_encode_warps.cuh -- K2w: the packing loop of compress_with_tree (comp.rs:422-447), one SUB-REGION PER WARP.
//
// hb_encode.cuh gives a CTA one region and its 32 warps 32 consecutive tiles per round: every round costs a CTA-wide
// barrier, an exchange of tile totals, a second scan, and every tile re-derives the last 32 bits of its predecessor from
// the predecessor's last 32 letters; words two lanes share are merged with shared-memory atomics into a zeroed staging
// area.  Here the histogram kernel counts every WARP's sub-region separately (148 x 32 = 4 736 of them), a tiny prepare
// kernel turns the counts into exact 64-bit bit offsets, and a warp then runs alone through its sub-region:
//   * no barrier, no exchange: the offset of the next tile is the running sum the warp already holds;
//   * the word two tiles share travels in a register (`carry`) from one tile to the next;
//   * the word two LANES share: a lane has 32 letters, so its bit string is at least 32 bits long and any stream word
//     holds bits of at most two lanes.  A lane stores every word it completes with a plain store, hands its last,
//     incomplete word to the next lane with one shuffle, and the receiver ORs it into the first word it stored.  No
//     shared-memory atomics, no zero-initialised staging;
//   * the lane-replicated table of packed (code << 16 | len) entries is addressed by one byte-permute (256 bytes per
//     letter, 128 used) + the table base the load instruction adds itself -- no 64 KiB of address padding.
// Codes of at most 15 bits (a tile then always fits its 512-word staging); longer codes take hb_encode.cuh.
//
// Algorithmic HBM bytes per launch: N + C (+ 1 KiB of counts per sub-region read by the prepare kernel).
#pragma once

#include "hb_common.cuh"
#include "hb_encode.cuh"

namespace hb {

constexpr int kEwWarps = 32;                                   // warps per CTA = sub-regions per CTA
constexpr int kEwThreads = kEwWarps * 32;
constexpr int kEwTile = 1024;                                  // letters per tile: 32 lanes x 32 consecutive letters
constexpr int kEwMaxBits = 15;                                 // longest code this kernel packs
constexpr int kEwStageWords = 516;                             // 31 + 1024 * 15 bits < 512 words; + slack
constexpr size_t kEwTableBytes = 256 * 256;                    // 256 bytes per letter: entry (b, lane) at b << 8 | lane << 2
constexpr size_t kEwSmemBytes = kEwTableBytes + 256 * sizeof(uint2) + static_cast<size_t>(kEwWarps) * kEwStageWords * 4 + 256;

// Sub-region bit totals from the per-sub-region histograms the histogram kernel wrote (sub_hist[s][b] = count of letter
// b in sub-region s).  One warp per sub-region.
// err: set to 1 when a letter that occurs has no code of 1..64 bits (the stream is then undefined).
__global__ void __launch_bounds__(256)
enc_prepare_kernel(const uint32_t *__restrict__ sub_hist, uint32_t n_sub,
                   const EncTable *__restrict__ table, unsigned long long *__restrict__ sub_bits, uint32_t *__restrict__ err) {
    const uint32_t s = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (s >= n_sub) return;
    unsigned long long acc = 0;
    uint32_t bad = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) {
        const uint32_t bsym = lane + 32 * k;
        const uint32_t c = sub_hist[static_cast<size_t>(s) * 256 + bsym];
        const uint32_t len = table->lo[bsym].y;
        acc += static_cast<unsigned long long>(c) * len;
        bad |= (c != 0 && len == 0) ? 1u : 0u;
    }
#pragma unroll
    for (int sft = 16; sft > 0; sft >>= 1) acc += __shfl_xor_sync(0xFFFFFFFFu, acc, sft);
    bad = __any_sync(0xFFFFFFFFu, bad);
    if (lane == 0) {
        sub_bits[s] = acc;
        if (bad) atomicOr(err, 1u);
    }
}

extern __shared__ __align__(16) uint8_t ew_smem[];

__global__ void __launch_bounds__(kEwThreads, 1)
encode_warps_kernel(const uint8_t *__restrict__ data, size_t n, const EncTable *__restrict__ table, uint32_t start_bit,
                    uint32_t *__restrict__ out32, const unsigned long long *__restrict__ sub_bits, uint32_t n_sub,
                    size_t sub_letters, unsigned long long *__restrict__ total_bits_out) {
    uint32_t *s_packed = reinterpret_cast<uint32_t *>(ew_smem);                       // [256][64]: 32 used per letter
    uint2 *s_tab = reinterpret_cast<uint2 *>(ew_smem + kEwTableBytes);                // [256] (code, len)
    uint32_t *s_stage_all = reinterpret_cast<uint32_t *>(ew_smem + kEwTableBytes + 256 * sizeof(uint2));
    unsigned long long *s_red = reinterpret_cast<unsigned long long *>(s_stage_all + kEwWarps * kEwStageWords);   // [32]

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 256 * 32; i += kEwThreads) s_packed[(i >> 5) * 64 + (i & 31)] = table->packed[i >> 5];
    for (int i = threadIdx.x; i < 256; i += kEwThreads) s_tab[i] = table->lo[i];

    // bit offset of this CTA's first sub-region: sum of all earlier sub-regions (every thread sums a strided part)
    const uint32_t first_sub = blockIdx.x * kEwWarps;
    {
        unsigned long long part = 0;
        for (uint32_t k = threadIdx.x; k < first_sub; k += kEwThreads) part += sub_bits[k];
#pragma unroll
        for (int sft = 16; sft > 0; sft >>= 1) part += __shfl_xor_sync(0xFFFFFFFFu, part, sft);
        if (lane == 0) s_red[warp] = part;
    }
    __syncthreads();
    unsigned long long running = start_bit;
    for (int k = 0; k < kEwWarps; k++) running += s_red[k];
    __syncthreads();                                         // s_red is read before anybody stages
    const uint32_t sub = first_sub + warp;
    if (sub >= n_sub) return;
    {
        // + the sub-regions of this CTA before mine
        unsigned long long mine = lane < warp ? sub_bits[first_sub + lane] : 0ull;
#pragma unroll
        for (int sft = 16; sft > 0; sft >>= 1) mine += __shfl_xor_sync(0xFFFFFFFFu, mine, sft);
        running += mine;
    }
    const size_t begin = static_cast<size_t>(sub) * sub_letters;
    if (begin >= n) return;
    const size_t end = min(n, begin + sub_letters);
    uint32_t *stage = s_stage_all + warp * kEwStageWords;
    const bool aligned32 = (reinterpret_cast<uintptr_t>(data) & 31) == 0;
    const uint32_t lane4 = static_cast<uint32_t>(lane) << 2;
    const uint8_t *tab_bytes = reinterpret_cast<const uint8_t *>(s_packed);

    // the word my first tile shares with the sub-region before me: its last (running % 32) bits, recomputed from the
    // 32 letters before `begin` (every code has at least one bit), left-aligned
    uint32_t carry = 0;
    if (begin > 0 && (running & 31)) {
        const uint2 e = s_tab[data[begin - 32 + lane]];
        uint32_t after = e.y;                                // suffix sum of the lengths of the letters after mine
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t o = __shfl_down_sync(0xFFFFFFFFu, after, d);
            if (lane + d < 32) after += o;
        }
        after -= e.y;
        const uint32_t low = e.y ? e.x >> (32 - e.y) : 0u;   // my code, right-aligned
        const uint32_t piece = after < 32 ? (low << after) : 0u;
        const uint32_t tail = __reduce_or_sync(0xFFFFFFFFu, piece);           // last 32 bits of the predecessor
        carry = tail << (32 - static_cast<uint32_t>(running & 31));
    }

    uint32_t raw[8];
    auto load_raw = [&](size_t at) {
        if (aligned32) {
            const u32x8 v = ld_stream_256(data + at);
#pragma unroll
            for (int j = 0; j < 8; j++) raw[j] = v.v[j];
        } else {
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const uint4 v = ld_stream_u4(reinterpret_cast<const uint4 *>(data + at) + h);
                raw[4 * h + 0] = v.x; raw[4 * h + 1] = v.y; raw[4 * h + 2] = v.z; raw[4 * h + 3] = v.w;
            }
        }
    };
    if (begin + kEwTile <= end) load_raw(begin + static_cast<size_t>(lane) * 32);

    for (size_t tile_base = begin; tile_base < end; tile_base += kEwTile) {
        const bool full = tile_base + kEwTile <= end;           // only the last tile of the whole input can be cut short
        const size_t lane_base = tile_base + static_cast<size_t>(lane) * 32;
        const uint32_t rr = static_cast<uint32_t>(running & 31);
        uint32_t tile_bits;
        if (full) {
            // ---- pass 1: 32 consecutive letters -> eight <= 60-bit pieces (four letters each) and their lengths
            uint32_t pv[16], plen[8], lane_bits = 0;
            constexpr uint32_t kCode = 0xFFFF0000u;
#pragma unroll
            for (int q = 0; q < 8; q++) {
                const uint32_t w = raw[q];
                // byte offset of entry (letter, lane): letter << 8 | lane << 2, built by one byte permute
                const uint32_t e0 = *reinterpret_cast<const uint32_t *>(tab_bytes + __byte_perm(w, lane4, 0x5504));
                const uint32_t e1 = *reinterpret_cast<const uint32_t *>(tab_bytes + __byte_perm(w, lane4, 0x5514));
                const uint32_t e2 = *reinterpret_cast<const uint32_t *>(tab_bytes + __byte_perm(w, lane4, 0x5524));
                const uint32_t e3 = *reinterpret_cast<const uint32_t *>(tab_bytes + __byte_perm(w, lane4, 0x5534));
                // entry = code << 16 | len: the wrapping funnel shift takes len from the low 5 bits
                const uint32_t v01 = (e0 & kCode) | __funnelshift_r(e1 & kCode, 0u, e0), l01 = (e0 + e1) & 0x3Fu;
                const uint32_t v23 = (e2 & kCode) | __funnelshift_r(e3 & kCode, 0u, e2), l23 = (e2 + e3) & 0x3Fu;
                pv[2 * q] = v01 | __funnelshift_rc(v23, 0u, l01);
                pv[2 * q + 1] = __funnelshift_rc(0u, v23, l01);
                plen[q] = l01 + l23;
                lane_bits += l01 + l23;
            }
            // `raw` is free: the next tile's letters, so the load latency hides behind the scan and the packing
            if (tile_base + 2 * kEwTile <= end) load_raw(lane_base + kEwTile);

            // ---- ONE warp scan places the lanes
            const uint32_t incl = warp_incl_scan(lane_bits);
            tile_bits = __shfl_sync(0xFFFFFFFFu, incl, 31);
            const uint32_t at = rr + incl - lane_bits;

            // ---- pass 2: 64-bit register packer; every completed word is a plain store (see the file header)
            uint32_t *ptr = stage + (at >> 5);
            uint32_t *const first = ptr;
            uint32_t fill = at & 31, hi = 0;
#pragma unroll
            for (int q = 0; q < 8; q++) {
                const uint32_t ph = pv[2 * q], pl = pv[2 * q + 1], L = plen[q];
                const uint32_t w0 = hi | (ph >> fill);
                const uint32_t w1 = __funnelshift_r(pl, ph, fill);
                const uint32_t w2 = __funnelshift_r(0u, pl, fill);
                const uint32_t nf = fill + L;
                if (nf >= 32) ptr[0] = w0;
                if (nf >= 64) ptr[1] = w1;
                hi = nf >= 64 ? w2 : (nf >= 32 ? w1 : w0);
                ptr += nf >> 5;
                fill = nf & 31;
            }
            // my incomplete last word goes to the next lane (lane 0 takes the carry of the previous tile), which ORs it
            // into the first word it stored; lane 31's becomes the next tile's carry
            uint32_t incoming = __shfl_up_sync(0xFFFFFFFFu, hi, 1);
            if (lane == 0) incoming = carry;
            carry = __shfl_sync(0xFFFFFFFFu, hi, 31);
            *first |= incoming;
        } else {
            // ---- the one cut-short tile at the end of the input: letter by letter into a zeroed staging area
            const uint32_t n_mine = lane_base >= end ? 0u : static_cast<uint32_t>(min(static_cast<size_t>(32), end - lane_base));
            const uint32_t lane_bits = enc_partial_bits<4>(data, lane_base, n_mine, s_tab);
            const uint32_t incl = warp_incl_scan(lane_bits);
            tile_bits = __shfl_sync(0xFFFFFFFFu, incl, 31);
            const uint32_t at = rr + incl - lane_bits;
            for (int k = lane; k < kEwStageWords; k += 32) stage[k] = 0;
            __syncwarp();
            if (lane == 0 && rr) atomicOr(&stage[0], carry);
            enc_partial_append<4>(data, lane_base, n_mine, s_tab, nullptr, stage + (at >> 5), at & 31);
            __syncwarp();
            carry = stage[(rr + tile_bits) >> 5];               // the stream's last, incomplete word
        }
        __syncwarp();

        // ---- copy out the complete words: coalesced big-endian 32-bit stores
        const unsigned long long w0g = running >> 5;
        const uint32_t n_full = (rr + tile_bits) >> 5;
        {
            const uint32_t *sp = stage + lane;
            uint32_t *gp = out32 + w0g + lane;
#pragma unroll 1
            for (uint32_t m = lane; m < n_full; m += 32, sp += 32, gp += 32) st_stream_u32(gp, bswap32(*sp));
        }
        running += tile_bits;
        __syncwarp();                                           // staging is reused by the next tile
    }

    if (end == n) {
        // the very end of the stream: pad bits are zero (comp.rs:446-447), write only the bytes that exist
        const uint32_t rem = static_cast<uint32_t>(running & 31);
        if (lane == 0) {
            if (rem) {
                uint8_t *dst8 = reinterpret_cast<uint8_t *>(out32 + (running >> 5));
                for (uint32_t k = 0; k < (rem + 7) / 8; k++) dst8[k] = static_cast<uint8_t>(carry >> (24 - 8 * k));
            }
            if (total_bits_out) *total_bits_out = running - start_bit;
        }
    }
}

}  // namespace hb
