// Host side of the codec: code construction with the reference's tie-breaking, the .crs2 header, and the
// small tables handed to the kernels. Pure C++ (no CUDA); microseconds per call.
#include <string.h>

#include <queue>
#include <vector>

#include "gh_internal.h"

namespace {

// comp(a, b) = freq[a] > freq[b]: a min-heap of symbol indices keyed by a frequency array that is mutated
// between pops and pushes -- the same shape as the reference's HuffNodeIndexGreater
// (include/canonical_huff_encoder.h:58-67). Using std::priority_queue over std::deque<int> itself means the
// libstdc++ heap order (which decides every tie) is inherited rather than imitated.
struct FreqGreater {
  const int64_t* f;
  explicit FreqGreater(const int64_t* freq) : f(freq) {}
  bool operator()(int a, int b) const { return f[a] > f[b]; }
};

inline void put_be32(uint8_t* p, uint32_t v) {
  p[0] = uint8_t(v >> 24);
  p[1] = uint8_t(v >> 16);
  p[2] = uint8_t(v >> 8);
  p[3] = uint8_t(v);
}
inline uint32_t get_be32(const uint8_t* p) {
  return (uint32_t(p[0]) << 24) | (uint32_t(p[1]) << 16) | (uint32_t(p[2]) << 8) | uint32_t(p[3]);
}

// Code lengths without a tree: pop the two lightest chains, deepen both by one, splice them, push the
// survivor back with the summed weight. Follows include/canonical_huff_encoder.cc:289-345.
int code_lengths(int64_t freq[GH_NSYM], uint32_t length[GH_NSYM], uint32_t* max_len) {
  int next_in_chain[GH_NSYM];
  // the reference keeps the heap in a std::deque; the heap algorithms -- and with them the order among equal weights
  // -- are the same for any random-access container, and a reserved vector is several times faster
  std::vector<int> storage;
  storage.reserve(GH_NSYM);
  std::priority_queue<int, std::vector<int>, FreqGreater> heap((FreqGreater(freq)), std::move(storage));
  for (int s = 0; s < GH_NSYM; ++s) {
    if (freq[s]) heap.push(s);
    next_in_chain[s] = -1;
    length[s] = 0;
  }
  for (int merges = int(heap.size()) - 1; merges > 0; --merges) {
    const int light = heap.top();
    heap.pop();
    const int heavy = heap.top();
    heap.pop();
    int s = heavy;
    for (; next_in_chain[s] != -1; s = next_in_chain[s]) length[s] += 1;
    next_in_chain[s] = light;  // tail of heavy's chain now continues into light's chain
    for (; s != -1; s = next_in_chain[s]) length[s] += 1;
    freq[heavy] += freq[light];
    heap.push(heavy);
  }
  uint32_t m = 0;
  for (int s = 0; s < GH_NSYM; ++s) m = length[s] > m ? length[s] : m;
  *max_len = m;
  if (m == 0) return GH_ERR_EMPTY;
  if (m > GH_MAX_CODE_LEN) return GH_ERR_TOO_LONG;
  return GH_OK;
}

// Canonical assignment, longest codes numerically smallest. Follows include/canonical_huff_encoder.cc:69-141.
int assign_codes(const uint32_t length[GH_NSYM], uint32_t max_len, gh_code* c) {
  uint32_t per_len[GH_MAX_CODE_LEN + 2] = {0};
  uint32_t next_code[GH_MAX_CODE_LEN + 2] = {0};
  uint32_t next_slot[GH_MAX_CODE_LEN + 2] = {0};
  memset(c, 0, sizeof(*c));
  c->max_len = max_len;
  for (int s = 0; s < GH_NSYM; ++s) {
    c->length[s] = length[s];
    c->symbol[s] = 0xFFFFFFFFu;
    if (length[s]) per_len[length[s]] += 1;
  }
  for (uint32_t len = 1; len <= max_len; ++len)
    if (per_len[len]) {
      c->min_len = len;
      break;
    }
  for (uint32_t len = 1; len <= max_len; ++len) c->start_pos[len] = c->start_pos[len - 1] + per_len[len - 1];
  c->first_code[max_len] = 0;
  for (uint32_t len = max_len; len-- > 1;) c->first_code[len] = (c->first_code[len + 1] + per_len[len + 1]) / 2;
  for (uint32_t len = 1; len <= max_len; ++len) {
    next_code[len] = c->first_code[len];
    next_slot[len] = c->start_pos[len];
  }
  for (uint32_t len = 1; len < c->min_len; ++len) c->first_code[len] = 1024;  // the reference's "never matches" mark
  for (int s = 0; s < GH_NSYM; ++s) {
    const uint32_t len = length[s];
    if (!len) continue;
    c->codeword[s] = next_code[len]++;
    c->symbol[next_slot[len]++] = uint32_t(s);
  }
  return GH_OK;
}

}  // namespace

extern "C" {

const char* gh_strerror(int status) {
  switch (status) {
    case GH_OK: return "ok";
    case GH_ERR_EMPTY: return "empty input (undefined in the reference)";
    case GH_ERR_TOO_LONG: return "code length > 32 (outside the reference's domain)";
    case GH_ERR_SPACE: return "output or workspace buffer too small";
    case GH_ERR_FORMAT: return "malformed header or stream";
    case GH_ERR_NO_EOF: return "stream ends before the end-of-encoding mark";
    case GH_ERR_ARG: return "bad argument";
    case GH_ERR_CUDA: return "CUDA failure (no device, no sm_100a image, or a launch error)";
    default: return "unknown status";
  }
}

int gh_build_code(const uint64_t hist256[256], gh_code* code) {
  if (!hist256 || !code) return GH_ERR_ARG;
  int64_t freq[GH_NSYM];
  uint32_t length[GH_NSYM], max_len = 0;
  for (int b = 0; b < 256; ++b) freq[b] = int64_t(hist256[b]);
  freq[GH_EOF_SYMBOL] = 1;  // include/encoder.h:128
  bool any = false;
  for (int b = 0; b < 256; ++b) any |= hist256[b] != 0;
  if (!any) return GH_ERR_EMPTY;
  const int rc = code_lengths(freq, length, &max_len);
  if (rc != GH_OK) return rc;
  return assign_codes(length, max_len, code);
}

size_t gh_header_bytes(const gh_code* code) { return 4 + 4 * size_t(GH_NSYM) + 8 + 8 * size_t(code->max_len); }

int gh_write_header(const gh_code* code, uint8_t* dst, size_t cap, size_t* written) {
  if (!code || !dst) return GH_ERR_ARG;
  if (code->max_len == 0 || code->max_len > GH_MAX_CODE_LEN) return GH_ERR_FORMAT;
  const size_t need = gh_header_bytes(code);
  if (cap < need) return GH_ERR_SPACE;
  uint8_t* p = dst;
  put_be32(p, GH_NSYM), p += 4;
  for (int i = 0; i < GH_NSYM; ++i, p += 4) put_be32(p, code->symbol[i]);
  put_be32(p, code->min_len), p += 4;
  put_be32(p, code->max_len), p += 4;
  for (uint32_t len = 1; len <= code->max_len; ++len, p += 8) {
    put_be32(p, code->start_pos[len]);
    put_be32(p + 4, code->first_code[len]);
  }
  if (written) *written = need;
  return GH_OK;
}

int gh_parse_header(const uint8_t* src, size_t n, gh_code* code, size_t* header_bytes) {
  if (!src || !code) return GH_ERR_ARG;
  memset(code, 0, sizeof(*code));
  const size_t fixed = 4 + 4 * size_t(GH_NSYM) + 8;
  if (n < fixed || get_be32(src) != GH_NSYM) return GH_ERR_FORMAT;
  const uint8_t* p = src + 4;
  for (int i = 0; i < GH_NSYM; ++i, p += 4) code->symbol[i] = get_be32(p);
  code->min_len = get_be32(p);
  code->max_len = get_be32(p + 4);
  p += 8;
  if (code->max_len == 0 || code->max_len > GH_MAX_CODE_LEN) return GH_ERR_FORMAT;
  if (code->min_len == 0 || code->min_len > code->max_len) return GH_ERR_FORMAT;
  if (n < gh_header_bytes(code)) return GH_ERR_FORMAT;
  for (uint32_t len = 1; len <= code->max_len; ++len, p += 8) {
    code->start_pos[len] = get_be32(p);
    code->first_code[len] = get_be32(p + 4);
  }
  // The tables are redundant: all of them follow from the code lengths (do_gen_encode, include/canonical_huff_encoder.cc:
  // 69-141). Only a header that is what the reference would have written for SOME set of lengths is accepted -- the
  // decoders index symbol_[start_pos_[len] + code - first_code_[len]] and trust these numbers.
  // (a) symbol_: the present symbols first, the rest is the reference's -1 fill
  uint32_t present = 0;
  while (present < GH_NSYM && code->symbol[present] < GH_NSYM) ++present;
  for (uint32_t k = present; k < GH_NSYM; ++k)
    if (code->symbol[k] != 0xFFFFFFFFu) return GH_ERR_FORMAT;
  if (present < 2) return GH_ERR_FORMAT;  // a byte value and the end mark at least
  // (b) start_pos_: where each length's symbols begin -> symbols per length
  uint32_t per_len[GH_MAX_CODE_LEN + 2] = {0};
  for (uint32_t len = 1; len <= code->max_len; ++len) {
    const uint32_t lo = code->start_pos[len];
    const uint32_t hi = len < code->max_len ? code->start_pos[len + 1] : present;
    if (lo > present || hi > present || hi < lo) return GH_ERR_FORMAT;
    if (len < code->min_len && hi != 0) return GH_ERR_FORMAT;
    per_len[len] = hi - lo;
  }
  if (code->start_pos[1] != 0 || per_len[code->min_len] == 0 || per_len[code->max_len] == 0) return GH_ERR_FORMAT;
  // (c) a Huffman code leaves no bit pattern unused: sum of 2^-len is exactly 1
  uint64_t kraft = 0;
  for (uint32_t len = 1; len <= code->max_len; ++len) kraft += uint64_t(per_len[len]) << (code->max_len - len);
  if (kraft != (1ull << code->max_len)) return GH_ERR_FORMAT;
  // (d) first_code_ as the reference derives it (entries below min_len hold its "never matches" mark: not compared)
  uint32_t expect = 0;
  for (uint32_t len = code->max_len; len >= code->min_len; --len) {
    if (code->first_code[len] != expect) return GH_ERR_FORMAT;
    expect = (expect + per_len[len]) / 2;
  }
  // (e) within a length the symbols ascend, and no symbol appears twice; this also derives the encoder-side view
  // (not stored in the file)
  for (uint32_t len = code->min_len; len <= code->max_len; ++len) {
    const uint32_t lo = code->start_pos[len];
    for (uint32_t k = lo; k < lo + per_len[len]; ++k) {
      const uint32_t s = code->symbol[k];
      if (code->length[s] != 0 || (k > lo && s <= code->symbol[k - 1])) return GH_ERR_FORMAT;
      code->length[s] = len;
      code->codeword[s] = code->first_code[len] + (k - lo);
    }
  }
  if (code->length[GH_EOF_SYMBOL] == 0) return GH_ERR_FORMAT;  // a stream that can never end
  if (header_bytes) *header_bytes = size_t(p - src);
  return GH_OK;
}

uint64_t gh_payload_bits(const gh_code* code, const uint64_t hist256[256], int with_eof) {
  uint64_t bits = with_eof ? code->length[GH_EOF_SYMBOL] : 0;
  for (int b = 0; b < 256; ++b) bits += hist256[b] * uint64_t(code->length[b]);
  return bits;
}

uint64_t gh_compress_bound(uint64_t n) { return 1040 + 8 * 32 + 4 * n + 32; }

}  // extern "C"

namespace gh {

int build_encode_table(const gh_code* code, EncodeTable* t) {
  memset(t, 0, sizeof(*t));
  if (code->max_len == 0 || code->max_len > GH_MAX_CODE_LEN) return GH_ERR_FORMAT;
  for (int s = 0; s < GH_NSYM; ++s) {
    if (code->length[s] > GH_MAX_CODE_LEN) return GH_ERR_TOO_LONG;
    t->codeword[s] = code->codeword[s];
    t->length[s] = uint8_t(code->length[s]);
  }
  return GH_OK;
}

int build_decode_tables(const gh_code* code, DecodeTables* t) {
  memset(t, 0, sizeof(*t));
  const uint32_t max_len = code->max_len;
  if (max_len == 0 || max_len > GH_MAX_CODE_LEN) return GH_ERR_FORMAT;
  if (code->min_len == 0 || code->min_len > max_len) return GH_ERR_FORMAT;
  t->min_len = code->min_len;
  t->max_len = max_len;
  for (int i = 0; i < GH_NSYM; ++i) t->symbol[i] = uint16_t(code->symbol[i] > GH_EOF_SYMBOL ? GH_EOF_SYMBOL : code->symbol[i]);
  for (uint32_t len = 1; len <= max_len; ++len) {
    // A first_code_ that no len-bit value can reach (the 1024 sentinel below min_len, or garbage in a
    // malformed header) becomes "greater than any window" in the left-justified form.
    const uint64_t fc = code->first_code[len];
    t->first_code_lj[len] = fc >= (1ull << len) ? 0xFFFFFFFFu : uint32_t(fc << (32 - len));
    t->start_pos[len] = code->start_pos[len];
  }
  return GH_OK;
}

}  // namespace gh
