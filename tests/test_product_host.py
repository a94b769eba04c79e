"""The product's host-side pieces that need no GPU, compiled with g++ straight from the headers:
Merlin transcript (host_transcript.hpp) against the published vector, the 64-bit host field code against the
32-bit carry-chain code that also runs on the device (fr.cuh / fq.cuh / host_fq64.hpp)."""
import os
import subprocess
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "lasso_b200", "csrc")

PROG = r'''
#include "host_transcript.hpp"
#include "ed25519.cuh"
#include "host_fq64.hpp"
#include <cstdio>
#include <random>
using namespace lb;
int main() {
  // 1. Merlin's published vector
  Transcript t("test protocol");
  t.append_message("some label", std::string("some data"));
  uint8_t o[32];
  t.challenge_bytes("challenge", o, 32);
  for (int i = 0; i < 32; i++) printf("%02x", o[i]);
  printf("\n");
  // 2. fast host Fr (64-bit limbs) == even/odd carry-chain multiplication (the device algorithm)
  std::mt19937_64 g(7);
  int bad = 0;
  fr_t a = fr_from_u64(g()), b = fr_from_u64(g());
  for (int i = 0; i < 20000; i++) {
    fr_t m1 = fr_mul(a, b), m2 = fr_mul_chain(a, b);
    if (!fr_eq(m1, m2)) bad++;
    a = fr_add(m1, b);
    b = fr_sub(m2, fr_from_u64(g()));
  }
  // fr_inv (host: 64-bit limbs, 4-bit window) == bitwise square-and-multiply on the carry-chain multiplication
  for (int i = 0; i < 300; i++) {
    fr_t x = fr_add(fr_mul(a, fr_from_u64(g())), fr_from_u64(g()));
    fr_t i1 = fr_inv(x), i2 = fr_inv_chain(x);
    if (!fr_eq(i1, i2) || !fr_eq(fr_mul(x, i1), fr_one())) bad++;
    a = x;
  }
  {
    fr_t one = fr_one(), m1 = fr_sub(fr_zero(), one);
    if (!fr_eq(fr_inv(one), one) || !fr_eq(fr_inv(m1), m1)) bad++;
  }
  // 3. host Fq64 normalisation == device-code normalisation + ark compression
  fq_t bx = {{0x8f25d51au, 0xc9562d60u, 0x9525a7b2u, 0x692cc760u, 0xfdd6dc5cu, 0xc0a4e231u, 0xcd6e53feu, 0x216936d3u}};
  fq_t by = {{0x66666658u, 0x66666666u, 0x66666666u, 0x66666666u, 0x66666666u, 0x66666666u, 0x66666666u, 0x66666666u}};
  pt_niels n = niels_from_affine(bx, by);
  pt_ext acc = pt_identity();
  for (int i = 0; i < 100; i++) {
    acc = pt_madd(pt_dbl(acc), n);
    fq_t x, y;
    pt_to_affine_canonical(acc, x, y);
    uint32_t c[8];
    pt_compress_canonical(x, y, c);
    uint32_t xyz[32];
    memcpy(xyz, acc.X.v, 32); memcpy(xyz + 8, acc.Y.v, 32); memcpy(xyz + 16, acc.Z.v, 32); memcpy(xyz + 24, acc.T.v, 32);
    uint8_t h[32];
    h64::compress_xyz(xyz, h);
    if (memcmp(h, c, 32)) bad++;
    // pair version (one inversion for two points) against the single one: this point and its double
    pt_ext dbl = pt_dbl(acc);
    uint32_t xyz2[32];
    memcpy(xyz2, dbl.X.v, 32); memcpy(xyz2 + 8, dbl.Y.v, 32); memcpy(xyz2 + 16, dbl.Z.v, 32); memcpy(xyz2 + 24, dbl.T.v, 32);
    uint8_t h2[32], pa[32], pb[32];
    h64::compress_xyz(xyz2, h2);
    h64::compress_xyz_pair(xyz, xyz2, pa, pb);
    if (memcmp(pa, h, 32) || memcmp(pb, h2, 32)) bad++;
  }
  printf("bad=%d\n", bad);
  return bad;
}
'''


def test_host_transcript_and_field_code():
    with tempfile.TemporaryDirectory() as d:
        src = os.path.join(d, "t.cpp")
        open(src, "w").write(PROG)
        exe = os.path.join(d, "t")
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-Wno-psabi", "-I", CSRC, src, "-o", exe])
        out = subprocess.run([exe], capture_output=True, text=True)
        lines = out.stdout.strip().splitlines()
        assert lines[0] == "d5a21972d0d5fe320c0d263fac7fffb8145aa640af6e9bca177c03c7efcf0615"
        assert lines[1] == "bad=0" and out.returncode == 0
