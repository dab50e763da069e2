"""Host model of the SIMD encode + validate of a 23-byte query (aindex_b200/csrc/query23.cuh::encode_validate23_rc):
the constants of the device code -- code bits (w ^ w >> 1) & 0x06060606, the packing multiplier 0x00820820 (top byte =
the four codes in reverse order), the PRMT tables "A.C." / "G.T." indexed by 2 * code, the byte gathers 0x0073 / 0x5410,
forward value = reverse_pairs(reversed codes) >> 18 -- restated in C with PRMT / BREV emulated, against the definition of
get_dna23_bitset (kmers.cpp:12-25) and reverseDNA (kmers.cpp:376-381).  Runs on the CPU: it pins the arithmetic, the
GPU parity tests pin the kernels."""
import os
import subprocess
import tempfile

SRC = r"""
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
static uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {  /* PTX prmt.b32, default mode */
    uint64_t ab = ((uint64_t)b << 32) | a; uint32_t r = 0;
    for (int i = 0; i < 4; i++) {
        uint32_t n = (sel >> (4 * i)) & 0xF, by = (uint32_t)(ab >> (8 * (n & 7))) & 0xFF;
        if (n & 8) by = (by & 0x80) ? 0xFF : 0;
        r |= by << (8 * i);
    }
    return r;
}
static uint64_t brevll(uint64_t x) { uint64_t r = 0; for (int i = 0; i < 64; i++) if (x >> i & 1) r |= 1ULL << (63 - i); return r; }
static uint64_t reverse_pairs64(uint64_t x) { x = brevll(x); return ((x & 0xAAAAAAAAAAAAAAAAULL) >> 1) | ((x & 0x5555555555555555ULL) << 1); }
static uint64_t revcomp23(uint64_t x) { return (~reverse_pairs64(x)) >> 18; }
static void model(const uint8_t *s, int *ok, uint64_t *u, uint64_t *r) {  /* s: 24 bytes, byte 23 is padding */
    uint32_t w[6], p[6], bad = 0;
    for (int j = 0; j < 6; j++) memcpy(&w[j], s + 4 * j, 4);
    for (int j = 0; j < 6; j++) {
        uint32_t x = (w[j] ^ (w[j] >> 1)) & 0x06060606u;
        p[j] = x * 0x00820820u;
        uint32_t sel = prmt(x + (x >> 4), 0, 0x4420);
        uint32_t d = prmt(0x00430041u, 0x00540047u, sel) ^ w[j];
        bad |= j == 5 ? (d & 0x00FFFFFFu) : d;
    }
    uint32_t p01 = prmt(p[0], p[1], 0x0073), p23 = prmt(p[2], p[3], 0x0073), p45 = prmt(p[4], p[5], 0x0073);
    uint32_t lo = prmt(p01, p23, 0x5410);
    uint64_t y = ((uint64_t)(p45 & 0x3FFFu) << 32) | lo;
    *r = y ^ 0x3FFFFFFFFFFFULL; *u = reverse_pairs64(y) >> 18; *ok = bad == 0;
}
static void definition(const uint8_t *s, int *ok, uint64_t *u) {
    *ok = 1; *u = 0;
    for (int j = 0; j < 23; j++) {
        int c;
        switch (s[j]) { case 'A': c = 0; break; case 'C': c = 1; break; case 'G': c = 2; break; case 'T': c = 3; break; default: c = 0; *ok = 0; }
        *u = (*u << 2) | (uint64_t)c;
    }
}
int main(void) {
    long nbad = 0; srand(1);
    for (long it = 0; it < 2000000; it++) {
        uint8_t s[24];
        for (int j = 0; j < 24; j++) s[j] = "ACGT"[rand() & 3];
        if (it % 3 == 1) s[rand() % 23] = (uint8_t)rand();
        if (it % 3 == 2) s[23] = (uint8_t)rand();                      /* the padding byte never matters */
        if (it % 7 == 3) for (int j = 0; j < 24; j++) s[j] = (uint8_t)rand();
        int ok1, ok2; uint64_t u1, r1, u2;
        model(s, &ok1, &u1, &r1); definition(s, &ok2, &u2);
        if (ok1 != ok2 || (ok2 && (u1 != u2 || r1 != revcomp23(u2)))) nbad++;
    }
    for (int pos = 0; pos < 23; pos++) for (int b = 0; b < 256; b++) {    /* every byte value at every position */
        uint8_t s[24]; memset(s, 'G', 24); s[pos] = (uint8_t)b;
        int ok1, ok2; uint64_t u1, r1, u2;
        model(s, &ok1, &u1, &r1); definition(s, &ok2, &u2);
        if (ok1 != ok2 || (ok2 && (u1 != u2 || r1 != revcomp23(u2)))) nbad++;
    }
    /* SURVEY 8(c) known answers */
    { int ok; uint64_t u, r; uint8_t s[24] = "ACGTACGTACGTACGTACGTACG"; model(s, &ok, &u, &r);
      if (!ok || u != 7450808207046ULL || r != 29803232828187ULL) nbad++; }
    printf("%ld\n", nbad);
    return nbad != 0;
}
"""


def test_encode_validate23_model():
    # the constants this model restates must be the ones the device code uses
    here = os.path.dirname(os.path.abspath(__file__))
    dev = open(os.path.join(here, "..", "aindex_b200", "csrc", "query23.cuh")).read()
    for token in ("0x06060606u", "0x00820820u", "0x4420u", "0x00430041u", "0x00540047u", "0x0073u", "0x5410u", "0x3FFFu",
                  "0x3FFFFFFFFFFFULL", "reverse_pairs64(y) >> 18"):
        assert token in dev, token
    with tempfile.TemporaryDirectory() as tmp:
        src, exe = os.path.join(tmp, "m.c"), os.path.join(tmp, "m")
        open(src, "w").write(SRC)
        subprocess.check_call(["gcc", "-O2", "-o", exe, src])
        out = subprocess.run([exe], stdout=subprocess.PIPE, text=True)
        assert out.returncode == 0 and out.stdout.strip() == "0", out.stdout
