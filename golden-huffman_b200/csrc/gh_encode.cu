// K2-K4 -- LUT gather, device-wide bit-offset scan and bit packing in ONE pass over the input
// (replaces reference include/canonical_huff_encoder.cc:245-285 encode_file/encode_each_byte and the
// per-bit writer utils/include/buffer.h:241-248,277-295).
//
// Algorithmic traffic: N bytes read + C bytes written (the scan is fused by decoupled look-back, so the
// input is not read a second time to learn the bit offsets, and every input byte is looked up ONCE).
//
// One persistent CTA of 1024 threads per SM, split into four independent GROUPS of 256 threads that share one
// lookup table and synchronise among themselves with named barriers only. A group takes 8 KiB TILES of input from
// an atomic ticket; a thread owns 32 consecutive bytes of a tile. Per tile k of a group:
//   1. gather + concatenate: per byte one PRMT (which forms the complete shared-memory address: the table has one
//      32-bit entry (code << 16 | len) per (byte value, lane) at 0x20000 + value * 256 + lane * 4, so the gathers of
//      a warp can never conflict) and one LDS; four codewords are concatenated into a 64-bit chunk by multiply-adds
//      with 2^len (one SHF each; FMA pipe), the chunk length is the sum of the entries' low halves;
//   2. warp-shuffle scan of the thread bit counts, the eight warp totals cross the first named barrier; the tile's
//      bit count is published at once as (AGGREGATE | bits);
//   3. staging: every chunk is OR-ed into one of the group's THREE zeroed staging buffers at its tile-relative bit
//      position with shared-memory atomics (neighbouring chunks share words); no global offset is needed for that;
//   4. warp 0 then resolves the start bit G of the group's PREVIOUS tile k-1 by a decoupled look-back (Merrill &
//      Garland) over 96 predecessors per L2 round trip, publishes (PREFIX | end bit) and draws the group's next tile;
//   5. meanwhile warps 1-7 copy out tile k-2, whose G has been known since the last iteration: global word (G/32 + i) =
//      funnel shift of staged words i-1, i by (G mod 32), byte-swapped to stream order, coalesced; the second barrier
//      of the iteration then frees that buffer (cleared; re-used by tile k+1).
//      A tile's count is thus published a whole tile period before anybody asks for it, and nobody waits for a
//      look-back. With look-back and copy-out right after staging, the groups ran in lock step with the slowest SM:
//      38-47 % of the warp samples sat at the barrier behind the look-back (profiles/rnd2_notes.md).
//      The tile's first word, when shared with the previous tile, is NOT stored: its bits go to head[tile] and
//      encode_stitch_kernel ORs them into the word the previous tile wrote -- every output word has exactly one
//      writer per kernel, no global atomics on the payload. The last tile adds the 1-padding (reference
//      include/canonical_huff_encoder.cc:255-257).
// Codewords longer than 16 bits cannot take part in step 1 (four of them do not fit a 64-bit chunk). Their table
// entries carry a flag that survives the length sum; a warp whose 1 KiB row of the tile contains one -- or the ragged
// end of the input, or the byte that the end mark follows -- handles that row codeword by codeword (count_slow /
// stage_slow). A Huffman code gives such lengths only to symbols rarer than 2^-16 or so. A tile whose bits do not fit
// one staging buffer (more than ~12 bits per byte on average) drains the group's pipeline, is staged across all three
// buffers and copied out at once (the "big tile" path).
#include <string.h>

#include "gh_common.cuh"

namespace gh {

constexpr int kEncGroupThreads = 256;
constexpr int kEncGroups = 4;
constexpr int kEncThreads = kEncGroupThreads * kEncGroups;  // one CTA per SM
constexpr int kEncGroupWarps = kEncGroupThreads / 32;
constexpr int kEncBytesPerThread = 32;
constexpr int kEncChunks = kEncBytesPerThread / 4;                    // 64-bit chunks of four codewords
constexpr int kEncTileBytes = kEncGroupThreads * kEncBytesPerThread;  // 8 KiB
constexpr int kEncRowBytes = 32 * kEncBytesPerThread;                 // a warp's row of a tile
constexpr int kEncLookDepth = 3;       // predecessors per lane and look-back round
constexpr u32 kEncLongFlag = 0x1000u;  // in the low half of a table entry: codeword longer than 16 bits

// Staging of a group: four zero words in front (staging ORs up to two words below a chunk's last word, copy-out
// reads word -1; 16-byte aligned clearing), then three buffers of 3068 words (a tile of up to 98048 bits, ~12 bits
// per byte; the last four words of a buffer stay zero). A tile of 8192 codewords of 32 bits + the end mark fits the
// three together.
constexpr int kEncStageFront = 4;
constexpr int kEncBufs = 3;
constexpr int kEncBufWords = 3068;
constexpr u32 kEncBufBits = u32(kEncBufWords - 4) * 32u;  // a tile with more bits than this takes the big-tile path
constexpr int kEncStageBytes = (kEncStageFront + kEncBufs * kEncBufWords) * 4;
static_assert(kEncBufs * kEncBufWords * 32 >= kEncTileBytes * 32 + 32 + 128, "a tile of 32-bit codewords must fit the three buffers");

// Shared memory. The replicated table must sit at a 64 KiB-aligned shared-window address so that a single PRMT can
// assemble an entry's address from (0x02, byte value, lane * 8); the groups' staging areas are laid out around it.
constexpr u32 kEncLutAddr = 0x20000u;
constexpr u32 kEncLutBytes = 256u * 256u;
struct EncGroupCtl {
  u32 wtot[kEncGroupWarps];  // warp totals of the tile being counted
  // the tiles staged in the group's buffers (group-uniform state lives here, not in every thread's registers)
  u64 s_tile[kEncBufs];
  u64 s_G[kEncBufs];
  u32 s_bits[kEncBufs];
  u32 next_tile;
};
struct EncLowSmem {  // at the start of dynamic shared memory
  EncGroupCtl ctl[kEncGroups];
  uint2 long_table[GH_NSYM + 1];  // (codeword, length) of every symbol, single copy: slow path and end mark
};

constexpr u64 kFlagMask = 3ull << 62;
constexpr u64 kFlagAggregate = 1ull << 62;  // value = bits of this tile only
constexpr u64 kFlagPrefix = 2ull << 62;     // value = global bit offset just after this tile

struct EncWorkspace {
  u64* tile_state;  // [ntiles]
  u32* head;        // [ntiles]
  u32* ticket;      // tile dispenser (tiles are handed out in the order groups ask for them)
};

__host__ __device__ inline u64 enc_num_tiles(u64 n) { return (n + kEncTileBytes - 1) / kEncTileBytes; }

// ---- named barrier of one group -----------------------------------------------------------------------------------
__device__ __forceinline__ void group_barrier(unsigned group) {
#ifdef GH_EMUL
  gh_emul::named_barrier(group + 1, kEncGroupThreads);
#else
  asm volatile("bar.sync %0, %1;" ::"r"(group + 1), "n"(kEncGroupThreads) : "memory");
#endif
}

// ---- table gather --------------------------------------------------------------------------------------------------
// entry of (byte value v, lane l): one 32-bit word (code << 16 | len [| long flag]) at kEncLutAddr + v * 256 + l * 4
#ifdef GH_EMUL
typedef const unsigned char* enc_lut_t;  // the table's base plus this lane's column
__device__ __forceinline__ u32 enc_lut_entry(enc_lut_t lut_lane, u32 word, int k) {
  return *reinterpret_cast<const u32*>(lut_lane + (((word >> (8 * k)) & 0xffu) << 8));
}
#else
typedef u32 enc_lut_t;  // kEncLutAddr | lane * 4
__device__ __forceinline__ u32 enc_lut_entry(enc_lut_t lut_lane, u32 word, int k) {
  // address = 0x00 0x02 <byte k of word> <lane * 4>: bytes 3, 2 and 0 from lut_lane, byte 1 from the input word
  const u32 addr = __byte_perm(word, lut_lane, 0x7604u | (u32(k) << 4));
  u32 v;
  asm("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
#endif

// ---- staging -------------------------------------------------------------------------------------------------------
// The staging buffer is addressed by its shared-window address (device) so that the ORs are RED.OR [reg + imm] without
// generic-pointer arithmetic.
#ifdef GH_EMUL
typedef u32* enc_stage_t;
__device__ __forceinline__ enc_stage_t enc_stage_handle(u32* p) { return p; }
__device__ __forceinline__ void enc_stage_or(enc_stage_t st, u32 word, int delta, u32 v) { atomicOr(st + int(word) + delta, v); }
__device__ __forceinline__ void enc_stage_or_nz(enc_stage_t st, u32 word, int delta, u32 v) {
  if (v) atomicOr(st + int(word) + delta, v);
}
__device__ __forceinline__ u32 enc_stage_ld(enc_stage_t st, u32 word, int delta) { return st[int(word) + delta]; }
#else
typedef u32 enc_stage_t;
__device__ __forceinline__ enc_stage_t enc_stage_handle(u32* p) { return u32(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void enc_stage_or(enc_stage_t st, u32 word, int delta, u32 v) {
  asm volatile("red.shared.or.b32 [%0], %1;" ::"r"(st + word * 4u + u32(delta * 4)), "r"(v) : "memory");
}
__device__ __forceinline__ void enc_stage_or_nz(enc_stage_t st, u32 word, int delta, u32 v) {
  asm volatile("{\n .reg .pred p;\n setp.ne.u32 p, %1, 0;\n @p red.shared.or.b32 [%0], %1;\n}" ::"r"(st + word * 4u + u32(delta * 4)), "r"(v)
               : "memory");
}
__device__ __forceinline__ u32 enc_stage_ld(enc_stage_t st, u32 word, int delta) {
  u32 v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(st + word * 4u + u32(delta * 4)) : "memory");
  return v;
}
#endif

// OR the low `len` bits of hi:lo (len in 1..64, higher bits zero) into the bit string so that they end just before
// bit `end` (= position + len). With W = end / 32 and e = end % 32, word W takes the value's last e bits at its top
// ((lo:0) >> e), word W-1 the 32 bits before them ((hi:lo) >> e) and word W-2 the rest (hi >> e): three funnel shifts
// by `end` itself (SHF takes its distance modulo 32), no other arithmetic. Words that receive nothing are OR-ed with
// zero (W, W-1) or skipped (W-2, rarely non-zero): the buffer has spare words on both sides.
__device__ __forceinline__ void stage_chunk(enc_stage_t stage, u32 end, u32 lo, u32 hi) {
  const u32 w = end >> 5;
  enc_stage_or(stage, w, 0, __funnelshift_r(0u, lo, end));
  enc_stage_or(stage, w, -1, __funnelshift_r(lo, hi, end));
  enc_stage_or_nz(stage, w, -2, __funnelshift_r(hi, 0u, end));
}

struct EncVec {
  u32 w[8];
};

__device__ __forceinline__ EncVec enc_load_vec(const uint8_t* p, bool aligned32) {
  const Unit8 u = ldg_unit(p, aligned32);
  EncVec v;
#pragma unroll
  for (int k = 0; k < 8; ++k) v.w[k] = u.w[k];
  return v;
}

// slow path: bit count / staging of the `cnt` bytes at p, codeword by codeword (bytes re-read from global memory)
__device__ __noinline__ u32 count_slow(const uint2* long_table, const uint8_t* p, int cnt) {
  u32 bits = 0;
  for (int k = 0; k < cnt; ++k) bits += long_table[p[k]].y;
  return bits;
}
__device__ __noinline__ u32 stage_slow(enc_stage_t stage, u32 pos, const uint2* long_table, const uint8_t* p, int cnt) {
  for (int k = 0; k < cnt; ++k) {
    const uint2 e = long_table[p[k]];
    pos += e.y;
    if (e.y) stage_chunk(stage, pos, e.x, 0u);
  }
  return pos;
}

// Decoupled look-back by one whole warp. The tile's own count was published before (AGGREGATE; tile 0: PREFIX);
// this adds up the predecessors' counts back to the nearest tile that already knows its start, publishes this
// tile's end bit and returns its start bit (all 32 lanes must call it together).
// `early` = the state of tile - 1 - lane, loaded by the caller some time before (0 = not loaded): it serves as a first
// round over 32 predecessors whose latency was hidden behind the caller's other work.
__device__ __forceinline__ u64 tile_start_lookback(const EncWorkspace& ws, u64 tile, u32 tile_bits, u64 start_bit, unsigned lane,
                                                   u64 early) {
  if (tile == 0) return start_bit;  // tile 0 published its PREFIX when it was counted
  u64 exclusive = 0;
  long long look = (long long)tile - 1;
  bool done = false;
  if (__all_sync(0xffffffffu, (early & kFlagMask) != 0)) {  // every one of the 32 is published
    const unsigned has_prefix = __ballot_sync(0xffffffffu, (early & kFlagMask) == kFlagPrefix);
    const unsigned first = unsigned(__ffs(int(has_prefix))) - 1u;
    exclusive = warp_sum64((has_prefix == 0 || lane <= first) ? (early & ~kFlagMask) : 0ull);
    done = has_prefix != 0;
    look -= 32;
  }
  while (!done) {
    // One round trip covers 32 x kEncLookDepth predecessors: lane l owns the consecutive tiles
    // look - l * kEncLookDepth - r (r = 0 nearest), loads all of them at once, folds them locally (sum of
    // aggregates up to and including its nearest PREFIX) and the warp then needs ONE ballot + ONE sum.
    const long long first_idx = look - (long long)lane * kEncLookDepth;
    u64 st[kEncLookDepth];
#pragma unroll
    for (int r = 0; r < kEncLookDepth; ++r) {
      const long long idx = first_idx - r;
      st[r] = idx >= 0 ? ld_volatile_u64(ws.tile_state + idx) : kFlagPrefix;  // virtual tiles before tile 0 add nothing
    }
    u64 local = 0;
    bool local_prefix = false;
#pragma unroll
    for (int r = 0; r < kEncLookDepth; ++r) {
      const long long idx = first_idx - r;
      while ((st[r] & kFlagMask) == 0) st[r] = ld_volatile_u64(ws.tile_state + idx);  // not published yet
      if (!local_prefix) local += st[r] & ~kFlagMask;
      local_prefix = local_prefix || (st[r] & kFlagMask) == kFlagPrefix;
    }
    const unsigned has_prefix = __ballot_sync(0xffffffffu, local_prefix);
    // lanes up to and including the nearest one that found a prefix contribute
    const unsigned first = unsigned(__ffs(int(has_prefix))) - 1u;
    const u64 contrib = (has_prefix == 0 || lane <= first) ? local : 0ull;
    exclusive += warp_sum64(contrib);
    done = has_prefix != 0;
    look -= 32 * kEncLookDepth;
  }
  if (lane == 0) st_volatile_u64(ws.tile_state + tile, kFlagPrefix | (exclusive + tile_bits));
  return exclusive;
}

// copy-out of one staged tile: `nthr` threads (this one is number `t`) store the words of the tile that starts at
// global bit G, reading the bit string staged at `stage`
__device__ __forceinline__ void enc_copy_out(enc_stage_t stage, u32 tile_bits, u64 tile, u64 G, bool last_tile, int append_eof,
                                             u32* __restrict__ out_words, u64 out_word_cap, u64* __restrict__ end_bit_out,
                                             u32* __restrict__ head, u32 t, u32 nthr) {
  const u32 phase = u32(G) & 31u;
  const u64 end_bit = G + tile_bits;
  const u64 word0 = G >> 5;
  const u32 nwords = (phase + tile_bits + 31u) >> 5;  // tile_bits > 0: every tile holds at least one codeword
  if (last_tile && t == 0 && end_bit_out) *end_bit_out = end_bit;
  u32* const dst = out_words + word0;
  const u64 room = out_word_cap > word0 ? out_word_cap - word0 : 0;
  const u32 lim = u64(nwords - 1u) < room ? nwords - 1u : u32(room);  // interior words: 1 .. nwords - 2
#ifdef GH_EMUL
  for (u32 i = t + 1u; i < lim; i += nthr)
    dst[i] = be32(__funnelshift_r(enc_stage_ld(stage, i, 0), enc_stage_ld(stage, i, -1), phase));
#else
  {
    // two words per thread and step (independent loads in flight), pointers advanced instead of re-derived
    u32 sa = stage + (t + 1u) * 4u;
    u32* gp = dst + (t + 1u);
    u32 i = t + 1u;
    for (; i + nthr < lim; i += 2u * nthr) {
      u32 a0, a1, b0, b1;
      asm("ld.shared.u32 %0, [%1];" : "=r"(a0) : "r"(sa));
      asm("ld.shared.u32 %0, [%1+-4];" : "=r"(a1) : "r"(sa));
      asm("ld.shared.u32 %0, [%1];" : "=r"(b0) : "r"(sa + nthr * 4u));
      asm("ld.shared.u32 %0, [%1+-4];" : "=r"(b1) : "r"(sa + nthr * 4u));
      gp[0] = be32(__funnelshift_r(a0, a1, phase));
      gp[nthr] = be32(__funnelshift_r(b0, b1, phase));
      sa += 2u * nthr * 4u;
      gp += 2u * nthr;
    }
    if (i < lim) {
      u32 a0, a1;
      asm("ld.shared.u32 %0, [%1];" : "=r"(a0) : "r"(sa));
      asm("ld.shared.u32 %0, [%1+-4];" : "=r"(a1) : "r"(sa));
      gp[0] = be32(__funnelshift_r(a0, a1, phase));
    }
  }
#endif
  // first and last word of the tile
  if (t < 2u && (t == 0 || nwords > 1u)) {
    const u32 i = t == 0 ? 0u : nwords - 1u;
    u32 v = __funnelshift_r(enc_stage_ld(stage, i, 0), enc_stage_ld(stage, i, -1), phase);
    if (append_eof && last_tile && i == nwords - 1u) {
      const u32 pad = u32((8 - (end_bit & 7)) & 7);  // flush_bits(): 1s up to the byte boundary
      if (pad) v |= ((1u << pad) - 1u) << (32u - (u32(end_bit & 31) + pad));
    }
    if (i == 0 && tile > 0 && phase != 0) head[tile] = v;  // shared with the previous tile: stitched later
    else if (u64(i) < room) dst[i] = be32(v);
  }
}

__device__ __forceinline__ void enc_clear(u32* buf, u32 tile_bits, u32 t) {
  uint4* const z = reinterpret_cast<uint4*>(buf);
  for (u32 i = t; i < (tile_bits >> 7) + 1u; i += kEncGroupThreads) z[i] = make_uint4(0u, 0u, 0u, 0u);
}

// kDeviceCode = false: the encode table arrives by value (constant bank), built on the host; true: the code was built on
// the device, table and placement are read from *dyn.
template <bool kDeviceCode>
__global__ void __launch_bounds__(kEncThreads, 1)
encode_kernel(const uint8_t* __restrict__ in, u64 n, const EncodeTable table_val, const gh_device_code* __restrict__ dyn,
              u64 start_bit, int append_eof, u32* __restrict__ out_words, u64 out_word_cap, u64* __restrict__ end_bit_out,
              EncWorkspace ws, u32 smem_bytes) {
  GH_DYNAMIC_SMEM(smem_raw);
  if (kDeviceCode) {
    // the code was built on the device: the payload follows the header build_code_kernel wrote, and is packed from the
    // 32-byte boundary below the header's end on (the decoder reads whole 32-byte sectors from there)
    if (dyn->status != u32(GH_OK)) return;
    const u32 hdr = dyn->header_bytes, base = hdr & ~31u;
    out_words += base / 4;
    out_word_cap = out_word_cap > base / 4 ? out_word_cap - base / 4 : 0;
    start_bit = u64(hdr - base) * 8;
  }
  const EncodeTable& table = kDeviceCode ? dyn->table : table_val;
  const unsigned group = threadIdx.x / kEncGroupThreads, tg = threadIdx.x % kEncGroupThreads;
  const unsigned lane = tg & 31, wg = tg >> 5;

  // ---- carve shared memory around the table's fixed address: staging areas below it first, then above it ----------
  EncLowSmem& low = *reinterpret_cast<EncLowSmem*>(smem_raw);
  constexpr u32 kLowBytes = (u32(sizeof(EncLowSmem)) + 15u) & ~15u;
#ifdef GH_EMUL
  const u32 smem_base = 0x400u;
#else
  const u32 smem_base = u32(__cvta_generic_to_shared(smem_raw));
#endif
  const u32 lut_off = kEncLutAddr - smem_base;
  const u32 below = (lut_off - kLowBytes) / u32(kEncStageBytes);  // groups that fit between the tables and the big table
  const u32 stage_off = group < below ? kLowBytes + group * u32(kEncStageBytes)
                                      : lut_off + kEncLutBytes + (group - below) * u32(kEncStageBytes);
#ifndef GH_EMUL
  if (smem_base + kLowBytes > kEncLutAddr || (smem_base & 15u) ||
      lut_off + kEncLutBytes + (u32(kEncGroups) - (below < u32(kEncGroups) ? below : u32(kEncGroups))) * u32(kEncStageBytes) > smem_bytes)
    __trap();  // the launch did not provide the window this layout needs
#endif
  (void)smem_bytes;
  unsigned char* const lut_ptr = smem_raw + lut_off;
  u32* const stage_base = reinterpret_cast<u32*>(smem_raw + stage_off);
  u32* const buf_ptr0 = stage_base + kEncStageFront;
  EncGroupCtl& ctl = low.ctl[group];

  // ---- tables (once per CTA) ---------------------------------------------------------------------------------------
  for (unsigned i = threadIdx.x; i < 256u * 32u; i += kEncThreads) {
    const unsigned sym = i >> 5, col = i & 31;
    const u32 len = table.length[sym];
    u32 e = 0;                                                  // byte value that does not occur
    if (len != 0 && len <= 16) e = (table.codeword[sym] << 16) | len;
    else if (len > 16) e = kEncLongFlag;                        // handled by the slow path
    *reinterpret_cast<u32*>(lut_ptr + sym * 256u + col * 4u) = e;
  }
  for (unsigned i = threadIdx.x; i < unsigned(GH_NSYM); i += kEncThreads)
    low.long_table[i] = make_uint2(table.codeword[i], u32(table.length[i]));
  for (unsigned i = tg; i < unsigned(kEncStageFront + kEncBufs * kEncBufWords); i += kEncGroupThreads) stage_base[i] = 0;
  if (tg == 0) ctl.next_tile = atomicAdd(ws.ticket, 1u);
  __syncthreads();
#ifdef GH_EMUL
  const enc_lut_t lut_lane = lut_ptr + lane * 4u;
#else
  const enc_lut_t lut_lane = kEncLutAddr | (lane * 4u);
#endif
  const uint2* const long_table = low.long_table;
  const u64 ntiles = enc_num_tiles(n);
  const bool aligned32 = (reinterpret_cast<uintptr_t>(in) & 31) == 0;
  const u32 eof_code = table.codeword[GH_EOF_SYMBOL];
  const u32 eof_len = append_eof ? u32(table.length[GH_EOF_SYMBOL]) : 0u;
  const u32 off = tg * kEncBytesPerThread;           // this thread's slice inside a tile
  const u32 row = (tg & ~31u) * kEncBytesPerThread;  // its warp's row

  // The group's pipeline: tile k is staged into buffer `cur` while tile k-1 (`a`, staged, start bit not resolved yet)
  // is resolved by warp 0 and tile k-2 (`b`, resolved) is copied out by warps 1-7. Iterations that cannot stage a tile
  // (no tile left; a big tile waiting for the pipeline to empty, or occupying all buffers) only advance the pipeline.
  u32 cur = 0, a_buf = 0, b_buf = 0;
  bool a_valid = false, b_valid = false;
  bool big_hold = false;  // a big tile is in the pipeline: nothing else may be staged
  const u32 total_groups = gridDim.x * u32(kEncGroups);

  u64 tile = ctl.next_tile;
  while (tile < ntiles || a_valid || b_valid) {
    bool staged = false, big = false;
    u32 c_lo[kEncChunks], c_hi[kEncChunks], c_end[kEncChunks];
    u32 bits = 0, pos = 0, tile_bits = 0;
    bool slow = false, row_live = false, owns_end = false;
    int cnt = 0;
    const uint8_t* tin = in;
    if (tile < ntiles && !big_hold) {
      const u64 tile_base = tile * kEncTileBytes;
      tin = in + tile_base;
      const u64 left = n - tile_base;
      const u32 rem = left < u64(kEncTileBytes) ? u32(left) : u32(kEncTileBytes);  // short only for the last tile
      const bool last_tile = (tile + 1 == ntiles);
      EncVec vec;
      if (off + kEncBytesPerThread <= rem) {
        vec = enc_load_vec(tin + off, aligned32);
      } else {
#pragma unroll
        for (int k = 0; k < 8; ++k) vec.w[k] = 0;
      }
      // ---- 1. gather + concatenate ---------------------------------------------------------------------------
      u32 flags = 0;
#pragma unroll
      for (int c = 0; c < kEncChunks; ++c) {
        const u32 word = vec.w[c];
        const u32 e0 = enc_lut_entry(lut_lane, word, 0);
        const u32 e1 = enc_lut_entry(lut_lane, word, 1);
        const u32 e2 = enc_lut_entry(lut_lane, word, 2);
        const u32 e3 = enc_lut_entry(lut_lane, word, 3);
        // 2^len of the three codewords that get shifted over: SHF takes its distance modulo 32 and len <= 16 sits in
        // the entry's low bits, so the entry itself is the shift operand
        const u32 p1 = __funnelshift_l(0u, 1u, e1), p2 = __funnelshift_l(0u, 1u, e2), p3 = __funnelshift_l(0u, 1u, e3);
        // acc = ((c0 * 2^l1 + c1) * 2^l2 * 2^l3) + (c2 * 2^l3 + c3): products on the FMA pipe; the additions cannot
        // carry because the products' low bits are zero
        const u32 a01 = __umulhi(e0, 1u << 16) * p1 + __umulhi(e1, 1u << 16);  // <= 32 bits
        const u32 a23 = __umulhi(e2, 1u << 16) * p3 + __umulhi(e3, 1u << 16);  // <= 32 bits
        const u64 t = u64(a01) * p2;                                            // <= 48 bits
        const u64 u = u64(u32(t)) * p3;
        c_lo[c] = u32(u) + a23;
        c_hi[c] = u32(t >> 32) * p3 + u32(u >> 32);
        const u32 l = (e0 + e1 + e2 + e3) & 0xffffu;  // lengths (and long-codeword flags) add up in the low half
        flags |= l;
        bits += l;
        c_end[c] = bits;  // end of this chunk relative to the thread's first bit
      }
      // rows that need the slow path: a long codeword, the ragged end, or the byte the end mark follows
      row_live = row < rem;
      const bool row_end = last_tile && row + kEncRowBytes >= rem;
      if (row_live) {
        slow = row_end || __any_sync(0xffffffffu, (flags & 0xf000u) != 0u);
        if (slow) {
          cnt = off >= rem ? 0 : (rem - off < u32(kEncBytesPerThread) ? int(rem - off) : kEncBytesPerThread);
          bits = count_slow(long_table, tin + off, cnt);
          owns_end = last_tile && cnt > 0 && off + u32(cnt) == rem;
          if (owns_end) bits += eof_len;
        }
      } else {
        bits = 0;
      }
      // ---- 2. scan -------------------------------------------------------------------------------------------
      const u32 incl = warp_inclusive_scan(bits, lane);
      if (lane == 31) ctl.wtot[wg] = incl;
      group_barrier(group);  // (a) warp totals; every clear of the previous iteration is done
      u32 wprefix;
      {
        // the eight warp totals, scanned by every warp for itself
        u32 v = lane < u32(kEncGroupWarps) ? ctl.wtot[lane] : 0u;
#pragma unroll
        for (int d = 1; d < kEncGroupWarps; d <<= 1) {
          const u32 up = __shfl_up_sync(0xffffffffu, v, d);
          if (lane >= unsigned(d)) v += up;
        }
        tile_bits = __shfl_sync(0xffffffffu, v, kEncGroupWarps - 1);
        wprefix = __shfl_sync(0xffffffffu, v, int(wg)) - __shfl_sync(0xffffffffu, incl, 31);
      }
      pos = wprefix + incl - bits;
      big = tile_bits > kEncBufBits;  // group-uniform
      // the tile's count is known to everybody a whole tile period before its own look-back runs
      if (tg == 0)
        st_volatile_u64(ws.tile_state + tile, tile > 0 ? (kFlagAggregate | u64(tile_bits)) : (kFlagPrefix | (start_bit + tile_bits)));
      // a big tile needs all three buffers: it waits (and is gathered again) until the pipeline is empty
      staged = !(big && (a_valid || b_valid));
      if (staged && big) cur = 0;
    }
    // warp 0 starts the look-back of tile k-1 now; the answer arrives while it stages its share of tile k
    // (and draws the group's next tile: the atomic's round trip is hidden the same way)
    u64 early = 0;
    u32 next_ticket = 0;
    if (wg == 0) {
      if (a_valid) {
        const u64 at = ctl.s_tile[a_buf];
        if (at > u64(lane)) early = ld_volatile_u64(ws.tile_state + (at - 1 - lane));
        else early = kFlagPrefix;  // virtual tiles before tile 0 add nothing
      }
      if (staged && lane == 0) next_ticket = atomicAdd(ws.ticket, 1u);
    }
    if (staged) {
      if (tg == 0) {  // read by the group after barrier (b)
        ctl.s_tile[cur] = tile;
        ctl.s_bits[cur] = tile_bits;
      }
      // ---- 3. staging at tile-relative positions ---------------------------------------------------------------
      const enc_stage_t stage = enc_stage_handle(buf_ptr0 + cur * kEncBufWords);
      if (!slow) {
        if (row_live) {
#pragma unroll
          for (int c = 0; c < kEncChunks; ++c) stage_chunk(stage, pos + c_end[c], c_lo[c], c_hi[c]);
        }
      } else {
        const u32 p2 = stage_slow(stage, pos, long_table, tin + off, cnt);
        if (owns_end && eof_len) stage_chunk(stage, p2 + eof_len, eof_code, 0u);
      }
    }
    // ---- 4. + 5. warp 0: the previous tile's start bit, the next tile's number; the others: copy-out of tile k-2 -----
    if (wg == 0) {
      if (a_valid) {
        const u64 g = tile_start_lookback(ws, ctl.s_tile[a_buf], ctl.s_bits[a_buf], start_bit, lane, early);
        if (lane == 0) ctl.s_G[a_buf] = g;
      }
      if (staged && lane == 0) {
        const u32 nt = next_ticket;
        ctl.next_tile = nt;
        // the tile somebody will draw about one tile period from now: start its way from DRAM to L2
        const u64 ahead = u64(nt) + total_groups;
        if (ahead + 1 < ntiles) prefetch_l2_bulk(in + ahead * kEncTileBytes, kEncTileBytes);
      }
    } else if (b_valid) {
      const u64 ptile = ctl.s_tile[b_buf];
      enc_copy_out(enc_stage_handle(buf_ptr0 + b_buf * kEncBufWords), ctl.s_bits[b_buf], ptile, ctl.s_G[b_buf], ptile + 1 == ntiles,
                   append_eof, out_words, out_word_cap, end_bit_out, ws.head, tg - 32u, kEncGroupThreads - 32u);
    }
    group_barrier(group);  // (b) tile k staged, G of tile k-1 and the next tile's number known, tile k-2 read
    if (b_valid) enc_clear(buf_ptr0 + b_buf * kEncBufWords, ctl.s_bits[b_buf], tg);
    b_valid = a_valid;
    b_buf = a_buf;
    a_valid = staged;
    if (staged) {
      a_buf = cur;
      cur = cur == u32(kEncBufs - 1) ? 0u : cur + 1u;
      big_hold = big;
      tile = ctl.next_tile;
    }
    if (big_hold && !a_valid && !b_valid) {  // the big tile has left: all three buffers are clean again
      big_hold = false;
      cur = 0;
    }
  }
}

// Second kernel: OR each tile's deferred head bits into the word its predecessor stored.
// Thread = tile. The first tile (in order) holding head bits for a given word merges the whole run.
__global__ void __launch_bounds__(256)
encode_stitch_kernel(u64 ntiles, const gh_device_code* __restrict__ dyn, u32* __restrict__ out_words, u64 out_word_cap,
                     EncWorkspace ws) {
  const u64 tile = u64(blockIdx.x) * blockDim.x + threadIdx.x;
  if (tile == 0 || tile >= ntiles) return;
  if (dyn) {  // same placement as encode_kernel
    if (dyn->status != u32(GH_OK)) return;
    const u32 base = dyn->header_bytes & ~31u;
    out_words += base / 4;
    out_word_cap = out_word_cap > base / 4 ? out_word_cap - base / 4 : 0;
  }
  const u64 g = ws.tile_state[tile - 1] & ~kFlagMask;  // global start bit of `tile`
  if ((g & 31) == 0) return;                            // starts on a word boundary: nothing deferred
  const u64 word = g >> 5;
  if (tile >= 2) {
    const u64 gp = ws.tile_state[tile - 2] & ~kFlagMask;  // start of the previous tile
    if ((gp >> 5) == word && (gp & 31) != 0) return;       // the previous tile deferred into this word too: it leads
  }
  u32 bits = ws.head[tile];
  for (u64 nx = tile + 1; nx < ntiles; ++nx) {  // only ever iterates for degenerate sub-word tiles
    const u64 gn = ws.tile_state[nx - 1] & ~kFlagMask;
    if ((gn >> 5) != word) break;
    bits |= ws.head[nx];
  }
  if (word < out_word_cap) out_words[word] |= be32(bits);
}

inline size_t enc_ws_bytes(u64 n) {
  const u64 nt = enc_num_tiles(n) + 1;
  return size_t(nt * 8 + ((nt * 4 + 7) / 8) * 8 + 256);
}

}  // namespace gh

extern "C" {

size_t gh_encode_workspace_bytes(uint64_t n) { return gh::enc_ws_bytes(n); }

uint64_t gh_encode_payload_capacity(uint64_t n, const gh_code* code, uint64_t start_bit) {
  const uint64_t max_len = code ? code->max_len : 32;
  const uint64_t bits = start_bit + n * max_len + 32 + 7;
  return ((bits + 127) / 128) * 16 + 16;
}

int gh_encode(const uint8_t* d_in, uint64_t n, const gh_code* code, uint64_t start_bit, int append_eof,
              uint8_t* d_payload, uint64_t payload_cap, uint64_t* d_end_bit, void* d_workspace,
              size_t workspace_bytes, void* stream) {
  if (code && payload_cap < gh_encode_payload_capacity(n, code, start_bit)) return GH_ERR_SPACE;
  return gh::encode_unchecked(d_in, n, code, start_bit, append_eof, d_payload, payload_cap, d_end_bit, d_workspace,
                              workspace_bytes, stream);
}

}  // extern "C"

namespace gh {

// the launches shared by both entry points: `table` by value (host-built code) or `dyn` (device-built code: table and
// placement of the payload behind the device-written header come from *dyn)
static int encode_launch(const uint8_t* d_in, u64 n, const EncodeTable& table, const gh_device_code* dyn, u64 start_bit,
                         int append_eof, uint8_t* d_payload, u64 payload_cap, u64* d_end_bit, uint8_t* ws_bytes_ptr, void* stream) {
  const u64 ntiles = enc_num_tiles(n);
  if (ntiles > 0x7ffffff0ull) return GH_ERR_ARG;
  EncWorkspace ws;
  uint8_t* p = ws_bytes_ptr;
  ws.tile_state = reinterpret_cast<u64*>(p);
  p += (ntiles + 1) * 8;
  ws.head = reinterpret_cast<u32*>(p);
  p += (((ntiles + 1) * 4 + 7) / 8) * 8;
  ws.ticket = reinterpret_cast<u32*>(p);
  const size_t used = size_t(p - ws_bytes_ptr) + 8;
  GH_CUDA_TRY(cudaMemsetAsync(ws_bytes_ptr, 0, used, (cudaStream_t)stream));

  // all of the SM's shared memory: the table's fixed window address decides the layout (see the kernel)
  int dev = 0, smem_max = 0;
  GH_CUDA_TRY(cudaGetDevice(&dev));
  GH_CUDA_TRY(cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  const u64 out_word_cap = payload_cap / 4;
  // persistent: one CTA per SM, four tile-taking groups each
  u64 blocks = u64(sm_count() > 0 ? sm_count() : 1);
  const u64 want = (ntiles + kEncGroups - 1) / kEncGroups;
  if (blocks > want) blocks = want;
  if (dyn) {
    GH_CUDA_TRY(cudaFuncSetAttribute(encode_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_max));
    GH_LAUNCH(encode_kernel<true>, unsigned(blocks), kEncThreads, size_t(smem_max), stream, d_in, (u64)n, table, dyn, (u64)start_bit,
              append_eof, reinterpret_cast<u32*>(d_payload), out_word_cap, d_end_bit, ws, u32(smem_max));
  } else {
    GH_CUDA_TRY(cudaFuncSetAttribute(encode_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_max));
    GH_LAUNCH(encode_kernel<false>, unsigned(blocks), kEncThreads, size_t(smem_max), stream, d_in, (u64)n, table, dyn, (u64)start_bit,
              append_eof, reinterpret_cast<u32*>(d_payload), out_word_cap, d_end_bit, ws, u32(smem_max));
  }
  int rc = check_launch();
  if (rc != GH_OK) return rc;
  if (ntiles > 1) {
    GH_LAUNCH(encode_stitch_kernel, unsigned((ntiles + 255) / 256), 256, 0, stream, ntiles, dyn,
              reinterpret_cast<u32*>(d_payload), out_word_cap, ws);
    rc = check_launch();
  }
  return rc;
}

int encode_unchecked(const uint8_t* d_in, uint64_t n, const gh_code* code, uint64_t start_bit, int append_eof,
                     uint8_t* d_payload, uint64_t payload_cap, uint64_t* d_end_bit, void* d_workspace,
                     size_t workspace_bytes, void* stream) {
  if (!code || !d_payload || !d_workspace) return GH_ERR_ARG;
  if (n == 0) return GH_ERR_EMPTY;
  if (!d_in) return GH_ERR_ARG;
  if ((reinterpret_cast<uintptr_t>(d_in) & 15) || (reinterpret_cast<uintptr_t>(d_payload) & 15) ||
      (reinterpret_cast<uintptr_t>(d_workspace) & 7))
    return GH_ERR_ARG;
  if (workspace_bytes < enc_ws_bytes(n)) return GH_ERR_SPACE;
  EncodeTable table;
  int rc = build_encode_table(code, &table);
  if (rc != GH_OK) return rc;
  return encode_launch(d_in, n, table, nullptr, start_bit, append_eof, d_payload, payload_cap, reinterpret_cast<u64*>(d_end_bit),
                       static_cast<uint8_t*>(d_workspace), stream);
}

int encode_with_device_code(const uint8_t* d_in, uint64_t n, const gh_device_code* d_code, uint8_t* d_image,
                            uint64_t image_cap, uint64_t* d_end_bit, void* d_workspace, size_t workspace_bytes, void* stream) {
  if (!d_code || !d_image || !d_workspace || !d_in) return GH_ERR_ARG;
  if (n == 0) return GH_ERR_EMPTY;
  if ((reinterpret_cast<uintptr_t>(d_in) & 15) || (reinterpret_cast<uintptr_t>(d_image) & 15) ||
      (reinterpret_cast<uintptr_t>(d_workspace) & 7))
    return GH_ERR_ARG;
  if (workspace_bytes < enc_ws_bytes(n)) return GH_ERR_SPACE;
  EncodeTable unused;
  memset(&unused, 0, sizeof(unused));
  return encode_launch(d_in, n, unused, d_code, 0, 1, d_image, image_cap / 4 * 4, reinterpret_cast<u64*>(d_end_bit),
                       static_cast<uint8_t*>(d_workspace), stream);
}

}  // namespace gh
