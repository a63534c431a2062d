// Wire format for keys, ciphertexts and plaintexts (host buffers).  The reference declares no serialisation at all (SURVEY 8f
// rank 4: "nothing exists"); multi-process sharding needs one, so this is the engine's own: a 64-byte little-endian header and the
// raw limb-major words the C ABI works on.
//
//   offset  0  char[8]  magic "FHEB200\0"
//           8  u32      version (1)
//          12  u32      kind     1 ciphertext, 2 public key, 3 secret key, 4 key-switching key (relinearisation / Galois), 5 plaintext
//          16  u32      n        ring degree
//          20  u32      limbs    RNS limbs per polynomial
//          24  u32      polys    polynomials in the payload (e.g. batch * 2 for ciphertexts, dnum * 2 for key-switching keys)
//          28  u32      flags    bit 0: NTT form
//          32  u32      galois_elt (0 unless a Galois key)
//          36  u32      reserved (0)
//          40  u64      moduli_hash   FNV-1a over the `limbs` moduli (little-endian bytes)
//          48  u64      payload_words = polys * limbs * n
//          56  u64      checksum      FNV-1a over the payload bytes
//          64  payload  uint64 little-endian [polys][limbs][n]
#include <cstdint>
#include <cstring>
#include "../../include/fhe_b200.h"

namespace fhe_b200 { void set_error(const char* fmt, ...); }
using fhe_b200::set_error;

namespace {
const char kMagic[8] = {'F', 'H', 'E', 'B', '2', '0', '0', '\0'};
uint64_t fnv1a(const void* p, size_t n, uint64_t h = 0xcbf29ce484222325ull) {
    const unsigned char* b = static_cast<const unsigned char*>(p);
    for (size_t i = 0; i < n; i++) { h ^= b[i]; h *= 0x100000001b3ull; }
    return h;
}
// sanity limits of the header fields (the engine's own: N <= 2^17, at most 4096 limbs): they also keep n * limbs * polys far from
// overflowing 64 bits, so a forged header cannot make the length check pass with a wrapped product
bool dims_ok(uint32_t n, uint32_t limbs, uint32_t polys) {
    return n >= 1 && n <= (1u << 17) && limbs >= 1 && limbs <= 4096 && polys >= 1 && polys <= (1u << 24);
}
void put32(uint8_t* p, uint32_t v) { for (int i = 0; i < 4; i++) p[i] = (uint8_t)(v >> (8 * i)); }
void put64(uint8_t* p, uint64_t v) { for (int i = 0; i < 8; i++) p[i] = (uint8_t)(v >> (8 * i)); }
uint32_t get32(const uint8_t* p) { uint32_t v = 0; for (int i = 0; i < 4; i++) v |= (uint32_t)p[i] << (8 * i); return v; }
uint64_t get64(const uint8_t* p) { uint64_t v = 0; for (int i = 0; i < 8; i++) v |= (uint64_t)p[i] << (8 * i); return v; }
}  // namespace

extern "C" size_t fhe_b200_wire_size(uint32_t n, uint32_t limbs, uint32_t polys) { return 64 + (size_t)n * limbs * polys * 8; }

extern "C" int fhe_b200_wire_pack(uint32_t kind, uint32_t n, uint32_t limbs, uint32_t polys, int ntt_form, uint32_t galois_elt,
                                  const uint64_t* h_moduli, const uint64_t* h_words, uint8_t* h_out) {
    if (!h_moduli || !h_words || !h_out || kind < 1 || kind > 5 || !dims_ok(n, limbs, polys)) {
        set_error("wire_pack: bad argument"); return FHE_B200_EINVAL;
    }
    const uint64_t words = (uint64_t)n * limbs * polys;
    memcpy(h_out, kMagic, 8);
    put32(h_out + 8, 1); put32(h_out + 12, kind); put32(h_out + 16, n); put32(h_out + 20, limbs); put32(h_out + 24, polys);
    put32(h_out + 28, ntt_form ? 1u : 0u); put32(h_out + 32, galois_elt); put32(h_out + 36, 0);
    uint8_t mb[8]; uint64_t mh = 0xcbf29ce484222325ull;
    for (uint32_t i = 0; i < limbs; i++) { put64(mb, h_moduli[i]); mh = fnv1a(mb, 8, mh); }
    put64(h_out + 40, mh); put64(h_out + 48, words);
    uint8_t* pay = h_out + 64;
    for (uint64_t i = 0; i < words; i++) put64(pay + 8 * i, h_words[i]);
    put64(h_out + 56, fnv1a(pay, words * 8));
    return 0;
}

/* h_words_out may be NULL to read the header only; h_moduli[n_moduli] (optional) is the caller's chain: the header's limb count must
 * equal n_moduli (checked BEFORE h_moduli is read: the header is untrusted), the chain's hash must match, and every payload word of
 * limb i must be below h_moduli[i] */
extern "C" int fhe_b200_wire_unpack(const uint8_t* h_in, size_t len, const uint64_t* h_moduli, uint32_t n_moduli, uint32_t* kind, uint32_t* n,
                                    uint32_t* limbs, uint32_t* polys, int* ntt_form, uint32_t* galois_elt, uint64_t* h_words_out) {
    if (!h_in || len < 64) { set_error("wire_unpack: buffer shorter than the header"); return FHE_B200_EINVAL; }
    if (memcmp(h_in, kMagic, 8) != 0) { set_error("wire_unpack: bad magic"); return FHE_B200_EINVAL; }
    if (get32(h_in + 8) != 1) { set_error("wire_unpack: unsupported version %u", get32(h_in + 8)); return FHE_B200_EINVAL; }
    const uint32_t k = get32(h_in + 12), nn = get32(h_in + 16), ll = get32(h_in + 20), pp = get32(h_in + 24);
    const uint64_t words = get64(h_in + 48);
    if (k < 1 || k > 5 || !dims_ok(nn, ll, pp) || words != (uint64_t)nn * ll * pp) { set_error("wire_unpack: inconsistent header"); return FHE_B200_EINVAL; }
    if (len != 64 + words * 8) { set_error("wire_unpack: %zu bytes given, header says %llu", len, (unsigned long long)(64 + words * 8)); return FHE_B200_EINVAL; }
    if (h_moduli) {
        if (ll != n_moduli) { set_error("wire_unpack: the object has %u limbs, the caller's chain %u", ll, n_moduli); return FHE_B200_ESTATE; }
        uint8_t mb[8]; uint64_t mh = 0xcbf29ce484222325ull;
        for (uint32_t i = 0; i < ll; i++) { put64(mb, h_moduli[i]); mh = fnv1a(mb, 8, mh); }
        if (mh != get64(h_in + 40)) { set_error("wire_unpack: the object was produced for a different modulus chain"); return FHE_B200_ESTATE; }
    }
    if (kind) *kind = k; if (n) *n = nn; if (limbs) *limbs = ll; if (polys) *polys = pp;
    if (ntt_form) *ntt_form = (int)(get32(h_in + 28) & 1u); if (galois_elt) *galois_elt = get32(h_in + 32);
    if (h_words_out) {
        if (fnv1a(h_in + 64, words * 8) != get64(h_in + 56)) { set_error("wire_unpack: payload checksum mismatch"); return FHE_B200_ESTATE; }
        for (uint64_t i = 0; i < words; i++) h_words_out[i] = get64(h_in + 64 + 8 * i);
        if (h_moduli)                                   // residues must be reduced: limb-major [polys][limbs][n]
            for (uint64_t i = 0; i < words; i++)
                if (h_words_out[i] >= h_moduli[(i / nn) % ll]) { set_error("wire_unpack: word %llu is not reduced modulo its limb", (unsigned long long)i); return FHE_B200_ESTATE; }
    }
    return 0;
}
