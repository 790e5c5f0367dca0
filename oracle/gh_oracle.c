/* TEST INFRASTRUCTURE ONLY -- see gh_oracle.h. Plain C11, single thread, no dependencies. */
#include "gh_oracle.h"

#include <string.h>

/* ---- a1 + a2 : include/encoder.h:123-129, 136-150 ------------------------------------------ */
void gho_histogram(const uint8_t* in, uint64_t n, int64_t freq[GHO_NSYM]) {
  for (int i = 0; i < GHO_NSYM - 1; i++) freq[i] = 0;
  freq[GHO_NSYM - 1] = 1; /* the end-of-encoding mark always occurs once */
  for (uint64_t i = 0; i < n; i++) freq[in[i]] += 1;
}

/* ---- the libstdc++ heap the reference's std::priority_queue<int, std::deque<int>, Cmp> runs on ----
 * Third-party arithmetic (not in the reference tree): libstdc++ 13.3.0, bits/stl_heap.h
 *   __push_heap :135-149, __adjust_heap :224-249, __pop_heap :254-267.
 * Comparator: include/canonical_huff_encoder.h:58-67, comp(a, b) = freq[a] > freq[b]  (a min-heap on freq).
 * The published algorithm is restated here because it decides every tie between equal frequencies. */
typedef struct {
  int a[GHO_NSYM];
  int n;
  const int64_t* f;
} gho_heap;

static void heap_sift_up(gho_heap* h, int hole, int top, int value) {
  int parent = (hole - 1) / 2;
  while (hole > top && h->f[h->a[parent]] > h->f[value]) {
    h->a[hole] = h->a[parent];
    hole = parent;
    parent = (hole - 1) / 2;
  }
  h->a[hole] = value;
}

static void heap_push(gho_heap* h, int value) { /* c.push_back(v); std::push_heap(...) */
  h->a[h->n++] = value;
  heap_sift_up(h, h->n - 1, 0, value);
}

static int heap_pop(gho_heap* h) { /* top(); std::pop_heap(...); c.pop_back() */
  int top = h->a[0];
  if (h->n > 1) {
    int len = h->n - 1; /* the heap shrinks first, the old last element is re-inserted from the root */
    int value = h->a[len];
    int hole = 0, child = 0;
    while (child < (len - 1) / 2) {
      child = 2 * (child + 1);
      if (h->f[h->a[child]] > h->f[h->a[child - 1]]) child--; /* comp(right, left) -> take left */
      h->a[hole] = h->a[child];
      hole = child;
    }
    if ((len & 1) == 0 && child == (len - 2) / 2) {
      child = 2 * (child + 1);
      h->a[hole] = h->a[child - 1];
      hole = child - 1;
    }
    heap_sift_up(h, hole, 0, value);
  }
  h->n--;
  return top;
}

/* ---- a3 : include/canonical_huff_encoder.cc:289-345 ---------------------------------------- */
int gho_encoding_lengths(int64_t freq[GHO_NSYM], uint32_t length[GHO_NSYM], uint32_t* max_len) {
  int group[GHO_NSYM];
  gho_heap h;
  h.n = 0;
  h.f = freq;
  for (int i = 0; i < GHO_NSYM; i++) { /* :299-304 push non-zero symbols in ascending order */
    if (freq[i]) heap_push(&h, i);
    group[i] = -1;
    length[i] = 0;
  }
  int times = h.n - 1;
  for (int t = 0; t < times; t++) { /* :308-334 */
    int top1 = heap_pop(&h);
    int top2 = heap_pop(&h);
    int index = top2;
    while (group[index] != -1) { /* walk top2's chain, one level deeper */
      length[index] += 1;
      index = group[index];
    }
    group[index] = top1; /* splice top1's chain on the tail */
    while (index != -1) {
      length[index] += 1;
      index = group[index];
    }
    freq[top2] += freq[top1]; /* top2 now stands for the merged internal node */
    heap_push(&h, top2);
  }
  uint32_t m = 0;
  for (int i = 0; i < GHO_NSYM; i++)
    if (length[i] > m) m = length[i];
  *max_len = m; /* :343 */
  if (m == 0) return GHO_ERR_EMPTY;
  if (m > 32) return GHO_ERR_TOO_LONG;
  return GHO_OK;
}

/* ---- a4 : include/canonical_huff_encoder.cc:69-141 ------------------------------------------ */
int gho_gen_code(const uint32_t length[GHO_NSYM], uint32_t max_len, gho_code* c) {
  if (max_len == 0) return GHO_ERR_EMPTY;
  if (max_len > 32) return GHO_ERR_TOO_LONG;
  uint32_t num[34], next_code[34], pos_copy[34];
  memset(c, 0, sizeof(*c));
  c->max_len = max_len;
  for (uint32_t i = 0; i <= max_len; i++) num[i] = 0;
  for (int i = 0; i < GHO_NSYM; i++) { /* :86-90 */
    num[length[i]] += 1;
    c->length[i] = length[i];
    c->symbol[i] = 0xFFFFFFFFu; /* symbol_[i] = -1 */
  }
  num[0] = 0;
  for (uint32_t i = 1; i <= max_len; i++) /* :93-98 */
    if (num[i] != 0) {
      c->min_len = i;
      break;
    }
  c->start_pos[0] = 0;
  for (uint32_t i = 1; i <= max_len; i++) c->start_pos[i] = num[i - 1] + c->start_pos[i - 1]; /* :104-105 */
  c->first_code[max_len] = 0; /* :109-114 */
  next_code[max_len] = 0;
  for (int i = (int)max_len - 1; i >= 1; i--) {
    c->first_code[i] = (c->first_code[i + 1] + num[i + 1]) / 2;
    next_code[i] = c->first_code[i];
  }
  for (uint32_t i = 1; i < c->min_len; i++) c->first_code[i] = 1024; /* :119-121 sentinel */
  for (uint32_t i = 0; i <= max_len; i++) pos_copy[i] = c->start_pos[i];
  for (int i = 0; i < GHO_NSYM; i++) { /* :127-133 */
    uint32_t len = length[i];
    if (len) {
      c->codeword[i] = next_code[len]++;
      c->symbol[pos_copy[len]++] = (uint32_t)i;
    }
  }
  return GHO_OK;
}

int gho_build_code(const uint64_t hist256[256], gho_code* code) {
  int64_t freq[GHO_NSYM];
  uint32_t length[GHO_NSYM], max_len = 0;
  for (int i = 0; i < 256; i++) freq[i] = (int64_t)hist256[i];
  freq[256] = 1;
  int rc = gho_encoding_lengths(freq, length, &max_len);
  if (rc != GHO_OK) return rc;
  return gho_gen_code(length, max_len, code);
}

/* ---- a5 : include/canonical_huff_encoder.cc:210-242, utils/include/buffer.h:255-268 ---------- */
static uint8_t* put_be32(uint8_t* p, uint32_t v) {
  p[0] = (uint8_t)(v >> 24);
  p[1] = (uint8_t)(v >> 16);
  p[2] = (uint8_t)(v >> 8);
  p[3] = (uint8_t)v;
  return p + 4;
}

size_t gho_header_bytes(const gho_code* c) { return 4 + 4 * GHO_NSYM + 8 + 8 * (size_t)c->max_len; }

size_t gho_write_header(const gho_code* c, uint8_t* dst) {
  uint8_t* p = dst;
  p = put_be32(p, GHO_NSYM);
  for (int i = 0; i < GHO_NSYM; i++) p = put_be32(p, c->symbol[i]);
  p = put_be32(p, c->min_len);
  p = put_be32(p, c->max_len);
  for (uint32_t i = 1; i <= c->max_len; i++) {
    p = put_be32(p, c->start_pos[i]);
    p = put_be32(p, c->first_code[i]);
  }
  return (size_t)(p - dst);
}

/* ---- a8 : include/canonical_huff_encoder.cc:349-374, utils/include/buffer.h:194-206 ---------- */
static uint32_t get_be32(const uint8_t* p) {
  return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3];
}

size_t gho_parse_header(const uint8_t* src, size_t n, gho_code* c) {
  memset(c, 0, sizeof(*c));
  if (n < 4 + 4 * GHO_NSYM + 8) return 0;
  if (get_be32(src) != GHO_NSYM) return 0;
  const uint8_t* p = src + 4;
  for (int i = 0; i < GHO_NSYM; i++, p += 4) c->symbol[i] = get_be32(p);
  c->min_len = get_be32(p);
  c->max_len = get_be32(p + 4);
  p += 8;
  if (c->max_len == 0 || c->max_len > 32) return 0;
  if (n < gho_header_bytes(c)) return 0;
  for (uint32_t i = 1; i <= c->max_len; i++, p += 8) {
    c->start_pos[i] = get_be32(p);
    c->first_code[i] = get_be32(p + 4);
  }
  /* the encoder-side tables are not in the file; rebuild them for convenience (not used by decode) */
  for (int i = 0; i < GHO_NSYM; i++) c->length[i] = 0;
  for (uint32_t len = 1; len <= c->max_len; len++) {
    uint32_t lo = c->start_pos[len];
    uint32_t hi = (len < c->max_len) ? c->start_pos[len + 1] : GHO_NSYM;
    for (uint32_t k = lo; k < hi && k < GHO_NSYM; k++) {
      uint32_t s = c->symbol[k];
      if (s >= GHO_NSYM) break;
      c->length[s] = len;
      c->codeword[s] = c->first_code[len] + (k - lo);
    }
  }
  return (size_t)(p - src);
}

uint64_t gho_payload_bits(const gho_code* c, const uint64_t hist256[256]) {
  uint64_t bits = c->length[GHO_EOF];
  for (int i = 0; i < 256; i++) bits += hist256[i] * c->length[i];
  return bits;
}

/* ---- a6 + a7 : include/canonical_huff_encoder.cc:245-285; utils/include/buffer.h:241-248,277-280,290-295 ----
 * write_bits emits bit len-1 .. 0 of the codeword, write_bit fills each byte from bit 7 down:
 * the stream is one MSB-first bit string. Restated with a 64-bit accumulator instead of one bit per call. */
int gho_encode_payload(const uint8_t* in, uint64_t n, const gho_code* c, uint8_t* dst, uint64_t cap,
                       uint64_t* payload_bytes) {
  uint64_t acc = 0; /* low `fill` bits are pending, oldest bit highest */
  unsigned fill = 0;
  uint64_t o = 0;
  for (uint64_t i = 0; i <= n; i++) {
    unsigned s = (i < n) ? in[i] : GHO_EOF; /* :255 the end mark goes last */
    unsigned len = c->length[s];
    uint64_t code = c->codeword[s];
    if (len == 0 || len > 32) return GHO_ERR_FORMAT;
    acc = (acc << len) | (code & ((len == 32) ? 0xFFFFFFFFull : ((1ull << len) - 1)));
    fill += len;
    while (fill >= 8) {
      if (o >= cap) return GHO_ERR_SPACE;
      dst[o++] = (uint8_t)(acc >> (fill - 8));
      fill -= 8;
    }
  }
  if (fill) { /* flush_bits(): pad the last byte with 1s (buffer.h:277-280) */
    unsigned pad = 8 - fill;
    if (o >= cap) return GHO_ERR_SPACE;
    dst[o++] = (uint8_t)((acc << pad) | ((1u << pad) - 1));
  }
  *payload_bytes = o;
  return GHO_OK;
}

/* ---- a9 : include/canonical_huff_encoder.cc:377-419 -----------------------------------------
 * One iteration per bit: v = (v<<1)|bit; ++len; if v >= first_code_[len] a symbol of that length ends here.
 * `v` is an int compared with an unsigned array in the reference, i.e. the comparison is unsigned. */
int gho_decode_payload(const uint8_t* payload, uint64_t nbytes, const gho_code* c, uint8_t* out, uint64_t cap,
                       uint64_t* n_out) {
  uint32_t v = 0;
  unsigned len = 0;
  uint64_t o = 0;
  for (uint64_t i = 0; i < nbytes; i++) {
    unsigned byte = payload[i];
    for (int b = 7; b >= 0; b--) {
      v = (v << 1) | ((byte >> b) & 1u);
      len++;
      if (len > c->max_len) return GHO_ERR_FORMAT; /* the reference would index past its tables */
      if (v >= c->first_code[len]) {
        uint32_t idx = c->start_pos[len] + v - c->first_code[len];
        if (idx >= GHO_NSYM) return GHO_ERR_FORMAT;
        uint32_t sym = c->symbol[idx];
        if (sym == GHO_EOF) {
          *n_out = o;
          return GHO_OK;
        }
        if (sym > GHO_EOF) return GHO_ERR_FORMAT;
        if (o >= cap) return GHO_ERR_SPACE;
        out[o++] = (uint8_t)sym;
        v = 0;
        len = 0;
      }
    }
  }
  *n_out = o;
  return GHO_ERR_NO_EOF; /* the reference would run off the end of the file */
}

/* ---- whole-file images : include/compressor.h:62-73, 87-92 ----------------------------------- */
uint64_t gho_compress_bound(uint64_t n) { return 1040 + 8 * 32 + 4 * n + 8; }

int gho_compress(const uint8_t* in, uint64_t n, uint8_t* out, uint64_t cap, uint64_t* out_bytes) {
  if (n == 0) return GHO_ERR_EMPTY;
  int64_t freq[GHO_NSYM];
  uint32_t length[GHO_NSYM], max_len = 0;
  gho_code code;
  gho_histogram(in, n, freq);                           /* encoder_.caculate_frequency() */
  int rc = gho_encoding_lengths(freq, length, &max_len); /* encoder_.gen_encode() */
  if (rc != GHO_OK) return rc;
  rc = gho_gen_code(length, max_len, &code);
  if (rc != GHO_OK) return rc;
  size_t hdr = gho_header_bytes(&code);
  if (cap < hdr) return GHO_ERR_SPACE;
  gho_write_header(&code, out); /* encoder_.write_encode_info() */
  uint64_t pbytes = 0;
  rc = gho_encode_payload(in, n, &code, out + hdr, cap - hdr, &pbytes); /* encoder_.encode_file() */
  if (rc != GHO_OK) return rc;
  *out_bytes = hdr + pbytes;
  return GHO_OK;
}

int gho_decompress(const uint8_t* in, uint64_t n, uint8_t* out, uint64_t cap, uint64_t* out_bytes) {
  gho_code code;
  size_t hdr = gho_parse_header(in, n, &code); /* decoder_.get_encode_info() */
  if (hdr == 0) return GHO_ERR_FORMAT;
  return gho_decode_payload(in + hdr, n - hdr, &code, out, cap, out_bytes); /* decoder_.decode_file() */
}
