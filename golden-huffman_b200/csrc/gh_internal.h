// Internal declarations shared by the host code and the CUDA translation units (not part of the C ABI).
#ifndef GH_INTERNAL_H_
#define GH_INTERNAL_H_

#include <stddef.h>
#include <stdint.h>

#include "gh_codec.h"

namespace gh {

// ---- decode tables handed to the device (built on the host from a gh_code) -------------------------
// Primary LUT indexed by the next kDecLutBits bits of the stream (MSB-first).
//   entry = (symbol << 6) | length      length in 1..kDecLutBits : codeword fully inside the window
//   entry = 0                           the codeword is longer than the window -> search first_code[]
// This is the device analogue of TableCanonicalHuffDecoder::lookup_table_
// (reference include/canonical_huff_encoder.cc:466-516) with the symbol folded in.
constexpr int kDecLutBits = 12;
constexpr int kDecLutSize = 1 << kDecLutBits;

struct DecodeTables {
  uint16_t lut[kDecLutSize];
  uint32_t first_code_lj[34];  // first_code_[len] << (32 - len), "left-justified" as in FastCanonicalHuffDecoder
                               // (reference include/canonical_huff_encoder.cc:437-438); [33] = 0 guard
  uint32_t start_pos[34];
  uint16_t symbol[GH_NSYM + 3];  // symbol_[] clamped to 0..256 (unused slots -> 256)
  uint32_t min_len, max_len;
};

// Fills `t` from `code` by replaying the reference's bit-serial rule
// (include/canonical_huff_encoder.cc:396-402) on every kDecLutBits-bit prefix.
int build_decode_tables(const gh_code* code, DecodeTables* t);

// ---- encode table: (codeword, length) per byte + the end mark, passed to kernels by value -----------
struct EncodeTable {
  uint32_t codeword[GH_NSYM];
  uint8_t length[GH_NSYM + 3];
};
int build_encode_table(const gh_code* code, EncodeTable* t);

// gh_decode with a non-zero first-codeword position (the payload of a .crs2 image starts 8-byte aligned, the
// kernels want 16: the image is decoded from the aligned address below with entry_bit = 64).
int decode_full(const uint8_t* d_payload, uint64_t payload_bytes, const gh_code* code, uint32_t entry_bit,
                uint8_t* d_out, uint64_t out_cap, uint64_t* n_out, void* d_workspace, size_t workspace_bytes,
                void* stream);

// gh_encode without the worst-case capacity precondition: stores are bounded by payload_cap instead.
int encode_unchecked(const uint8_t* d_in, uint64_t n, const gh_code* code, uint64_t start_bit, int append_eof,
                     uint8_t* d_payload, uint64_t payload_cap, uint64_t* d_end_bit, void* d_workspace,
                     size_t workspace_bytes, void* stream);

}  // namespace gh
#endif
