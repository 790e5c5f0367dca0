// Code construction ON THE DEVICE: the 257-symbol code is built from the histogram by one
// warp, bit-for-bit as the reference builds it, so that compress can run histogram -> code -> header -> packing on one
// stream without the host in between. The host builder gh_build_code (gh_host.cc) stays: it is the checker in the
// tests and serves every caller that starts from a host histogram.
//
//   1. Encoder::do_init (reference include/encoder.h:123-129): the end mark's count is 1.
//   2. CanonicalHuffEncoder::get_encoding_length (include/canonical_huff_encoder.cc:289-345): a min-heap of symbol
//      indices keyed by the live frequency array; pop two, deepen both chains by one, splice, push the survivor with
//      the summed weight. Which of two equal frequencies is popped first is decided by libstdc++'s heap
//      (std::priority_queue<int, std::deque<int>, Cmp>: bits/stl_heap.h 13.3.0, __push_heap :135-149, __adjust_heap
//      :224-249, __pop_heap :254-267), so that heap is restated here step by step. Lane 0 runs it on (key, index)
//      pairs in shared memory -- a key cannot change while its index is inside the heap, so carrying it along is the
//      same as the reference's look-up through the index. The chain walks become set operations: every symbol carries
//      the head of its chain, and "deepen the two chains and splice them" is `length += 1, head = survivor` for every
//      symbol whose head is one of the two -- all 32 lanes, nine symbols each, in registers.
//   3. CanonicalHuffEncoder::do_gen_encode (:69-141): counts per length, start_pos_, first_code_ (longest codes
//      numerically smallest), the 1024 sentinel below min_len, codewords and symbol_ in ascending symbol order.
//   4. write_encode_info (:210-242): the big-endian header, written straight to its place in the output image.
#include "gh_common.cuh"

namespace gh {

struct BuildHeapEntry {
  long long key;
  u32 idx;
  u32 pad;
};

struct BuildSmem {
  BuildHeapEntry heap[GH_NSYM + 1];
  long long freq[GH_NSYM + 1];
  u32 length[GH_NSYM + 3];
  u32 codeword[GH_NSYM + 3];
  u32 symbol[GH_NSYM + 3];
  u32 per_len[GH_MAX_CODE_LEN + 2];
  u32 start_pos[GH_MAX_CODE_LEN + 2];
  u32 first_code[GH_MAX_CODE_LEN + 2];
  u32 next_code[GH_MAX_CODE_LEN + 2];
  u32 next_slot[GH_MAX_CODE_LEN + 2];
  int heap_n;
};

// libstdc++ __push_heap: the value climbs from `hole` while its parent compares greater (comp(parent, value) =
// freq[parent] > freq[value])
__device__ __forceinline__ void build_sift_up(BuildSmem& s, int hole, long long key, u32 idx) {
  while (hole > 0) {
    const int parent = (hole - 1) / 2;
    const BuildHeapEntry p = s.heap[parent];
    if (!(p.key > key)) break;
    s.heap[hole] = p;
    hole = parent;
  }
  BuildHeapEntry e;
  e.key = key, e.idx = idx, e.pad = 0;
  s.heap[hole] = e;
}

__device__ __forceinline__ void build_push(BuildSmem& s, long long key, u32 idx) {  // c.push_back(v); std::push_heap
  build_sift_up(s, s.heap_n, key, idx);
  s.heap_n += 1;
}

// top(); std::pop_heap (the heap shrinks first; the old last element re-enters from the root: the hole sinks to a
// leaf, always towards the right child unless comp(right, left), then the element climbs back); c.pop_back()
__device__ __forceinline__ BuildHeapEntry build_pop(BuildSmem& s) {
  const BuildHeapEntry top = s.heap[0];
  const int len = s.heap_n - 1;
  if (len > 0) {
    const BuildHeapEntry value = s.heap[len];
    int hole = 0, child = 0;
    while (child < (len - 1) / 2) {
      child = 2 * (child + 1);
      const BuildHeapEntry r = s.heap[child], l = s.heap[child - 1];
      if (r.key > l.key) {  // comp(right, left): take the left one
        s.heap[hole] = l;
        child -= 1;
      } else {
        s.heap[hole] = r;
      }
      hole = child;
    }
    if ((len & 1) == 0 && child == (len - 2) / 2) {
      child = 2 * (child + 1);
      s.heap[hole] = s.heap[child - 1];
      hole = child - 1;
    }
    build_sift_up(s, hole, value.key, value.idx);
  }
  s.heap_n = len;
  return top;
}

constexpr int kBuildPerLane = (GH_NSYM + 31) / 32;  // 9 symbols per lane: symbol = lane + 32 * k

// One warp. d_hists: n_hists histograms of 256 u64 counters each (summed: the shards of a multi-GPU input).
// d_header (optional): where the header's bytes go (the front of the output image).
__global__ void __launch_bounds__(32)
build_code_kernel(const u64* __restrict__ d_hists, int n_hists, gh_device_code* __restrict__ out, uint8_t* __restrict__ d_header) {
  __shared__ BuildSmem s;
  const unsigned lane = threadIdx.x;
  // ---- 1. frequencies ---------------------------------------------------------------------------------------------
  u64 total = 0;
  for (unsigned b = lane; b < 256; b += 32) {
    u64 f = 0;
    for (int h = 0; h < n_hists; ++h) f += d_hists[size_t(h) * 256 + b];
    s.freq[b] = (long long)f;
    total += f;
  }
  total = warp_sum64(total);
  if (lane == 0) {
    s.freq[GH_EOF_SYMBOL] = 1;
    s.heap_n = 0;
  }
  for (unsigned i = lane; i < unsigned(GH_NSYM + 3); i += 32) s.length[i] = 0, s.codeword[i] = 0, s.symbol[i] = 0xFFFFFFFFu;
  for (unsigned i = lane; i < unsigned(GH_MAX_CODE_LEN + 2); i += 32) s.per_len[i] = 0, s.start_pos[i] = 0, s.first_code[i] = 0;
  __syncwarp();
  u32 status = total == 0 ? u32(GH_ERR_EMPTY) : u32(GH_OK);
  // ---- 2. code lengths -----------------------------------------------------------------------------------------------
  u32 head[kBuildPerLane], len[kBuildPerLane];
#pragma unroll
  for (int k = 0; k < kBuildPerLane; ++k) head[k] = lane + 32u * k, len[k] = 0;
  if (lane == 0) {
    for (int sym = 0; sym < GH_NSYM; ++sym)  // ascending symbol order, non-zero counts only (:299-304)
      if (s.freq[sym]) build_push(s, s.freq[sym], u32(sym));
  }
  __syncwarp();
  int merges = __shfl_sync(0xffffffffu, s.heap_n, 0) - 1;  // lane 0's view: it changes heap_n right away below
  for (; merges > 0 && status == u32(GH_OK); --merges) {
    u32 light = 0, heavy = 0;
    if (lane == 0) {
      const BuildHeapEntry a = build_pop(s);
      const BuildHeapEntry b = build_pop(s);
      light = a.idx, heavy = b.idx;
      const long long sum = a.key + b.key;
      s.freq[heavy] = sum;  // the survivor stands for the merged node
      build_push(s, sum, heavy);
    }
    light = __shfl_sync(0xffffffffu, light, 0);
    heavy = __shfl_sync(0xffffffffu, heavy, 0);
#pragma unroll
    for (int k = 0; k < kBuildPerLane; ++k) {
      const bool in = head[k] == light || head[k] == heavy;
      len[k] += in ? 1u : 0u;
      head[k] = in ? heavy : head[k];
    }
  }
  u32 max_len = 0;
#pragma unroll
  for (int k = 0; k < kBuildPerLane; ++k) {
    const unsigned sym = lane + 32u * k;
    if (sym < unsigned(GH_NSYM)) {
      s.length[sym] = len[k];
      max_len = len[k] > max_len ? len[k] : max_len;
      if (len[k] && len[k] <= u32(GH_MAX_CODE_LEN)) atomicAdd(&s.per_len[len[k]], 1u);
    }
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    const u32 o = __shfl_xor_sync(0xffffffffu, max_len, d);
    max_len = o > max_len ? o : max_len;
  }
  if (status == u32(GH_OK) && max_len == 0) status = u32(GH_ERR_EMPTY);
  if (status == u32(GH_OK) && max_len > u32(GH_MAX_CODE_LEN)) status = u32(GH_ERR_TOO_LONG);
  __syncwarp();
  // ---- 3. canonical assignment -----------------------------------------------------------------------------------------
  u32 min_len = 0;
  if (status == u32(GH_OK)) {
    if (lane == 0) {
      for (u32 l = 1; l <= max_len; ++l)
        if (s.per_len[l]) {
          min_len = l;
          break;
        }
      for (u32 l = 1; l <= max_len; ++l) s.start_pos[l] = s.start_pos[l - 1] + s.per_len[l - 1];
      s.first_code[max_len] = 0;
      for (u32 l = max_len; l-- > 1;) s.first_code[l] = (s.first_code[l + 1] + s.per_len[l + 1]) / 2;
      for (u32 l = 1; l <= max_len; ++l) s.next_code[l] = s.first_code[l], s.next_slot[l] = s.start_pos[l];
      for (u32 l = 1; l < min_len; ++l) s.first_code[l] = 1024;  // the reference's "never matches" mark
      for (int sym = 0; sym < GH_NSYM; ++sym) {
        const u32 l = s.length[sym];
        if (!l) continue;
        s.codeword[sym] = s.next_code[l]++;
        s.symbol[s.next_slot[l]++] = u32(sym);
      }
    }
    min_len = __shfl_sync(0xffffffffu, min_len, 0);
  }
  __syncwarp();
  // ---- 4. results: the code, the kernels' encode table, the payload size, the header ---------------------------------------
  for (unsigned sym = lane; sym < unsigned(GH_NSYM); sym += 32) {
    const u32 l = status == u32(GH_OK) ? s.length[sym] : 0u;
    out->code.length[sym] = l;
    out->code.codeword[sym] = s.codeword[sym];
    out->code.symbol[sym] = s.symbol[sym];
    out->table.codeword[sym] = s.codeword[sym];
    out->table.length[sym] = uint8_t(l);
  }
  // payload bits: sum len[b] * count[b] over the ORIGINAL counts (s.freq was overwritten by the merges) + the end mark
  u64 bits = 0;
  for (unsigned b = lane; b < 256; b += 32) {
    u64 f = 0;
    for (int h = 0; h < n_hists; ++h) f += d_hists[size_t(h) * 256 + b];
    bits += f * u64(status == u32(GH_OK) ? s.length[b] : 0u);
  }
  bits = warp_sum64(bits);
  for (unsigned l = lane; l < 33; l += 32) {
    out->code.start_pos[l] = l <= max_len && status == u32(GH_OK) ? s.start_pos[l] : 0u;
    out->code.first_code[l] = l <= max_len && status == u32(GH_OK) ? s.first_code[l] : 0u;
  }
  const u32 header_bytes = 4u + 4u * u32(GH_NSYM) + 8u + 8u * max_len;
  if (lane == 0) {
    out->code.min_len = min_len;
    out->code.max_len = max_len;
    out->status = status;
    out->header_bytes = status == u32(GH_OK) ? header_bytes : 0u;
    out->payload_bits = status == u32(GH_OK) ? bits + s.length[GH_EOF_SYMBOL] : 0ull;
    out->total_symbols = total;
  }
  if (d_header && status == u32(GH_OK)) {
    u32* const hw = reinterpret_cast<u32*>(d_header);  // big-endian 32-bit fields, the image is 16-byte aligned
    const u32 nwords = header_bytes / 4;
    for (u32 w = lane; w < nwords; w += 32) {
      u32 v;
      if (w == 0) v = u32(GH_NSYM);
      else if (w <= u32(GH_NSYM)) v = s.symbol[w - 1];
      else if (w == u32(GH_NSYM) + 1) v = min_len;
      else if (w == u32(GH_NSYM) + 2) v = max_len;
      else {
        const u32 k = w - (u32(GH_NSYM) + 3);  // pairs (start_pos[i], first_code[i]) for i = 1 .. max_len
        const u32 i = k / 2 + 1;
        v = (k & 1) ? s.first_code[i] : s.start_pos[i];
      }
      hw[w] = be32(v);
    }
  }
}

}  // namespace gh

extern "C" {

int gh_build_code_device(const uint64_t* d_hists, int n_hists, gh_device_code* d_code, uint8_t* d_header, void* stream) {
  using namespace gh;
  if (!d_hists || !d_code || n_hists < 1) return GH_ERR_ARG;
  if ((reinterpret_cast<uintptr_t>(d_code) & 7) || (reinterpret_cast<uintptr_t>(d_header) & 3)) return GH_ERR_ARG;
  GH_LAUNCH(build_code_kernel, 1, 32, 0, stream, reinterpret_cast<const u64*>(d_hists), n_hists, d_code, d_header);
  return check_launch();
}

}  // extern "C"
