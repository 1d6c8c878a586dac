#!/usr/bin/env python
"""CPU study (no GPU): algorithmic error of operand-split schemes for fp32-grade GEMMs on TF32/BF16 tensor cores.
Operands are rounded exactly as the tensor core would see them; products and sums are exact (float64), so the numbers
isolate the SPLIT error from accumulation effects (those are measured on the device, conv_gemm.cu 'r01 finding 2').

  3xTF32        a_hi*b_hi + a_lo*b_hi + a_hi*b_lo, all operands tf32            (shipped: 3 TF32 MMAs per product)
  tf32+2xbf16   a_hi*b_hi in tf32, the two cross terms with bf16 operands       (candidate: 1 TF32 + 2 BF16 MMAs = 2/3 the
                                                                                 tensor-pipe time, same smem/TMEM bytes)
  1xTF32        a_hi*b_hi only
Output: max |err| / max |out| over a [256, K] x [K, 64] product with ReLU-like activations and He-scaled weights."""
import numpy as np


def round_mantissa(x, bits):
    """fp32 -> `bits` explicit mantissa bits, round to nearest, ties away from zero (cvt.rna)."""
    u = np.asarray(x, np.float32).view(np.uint32).astype(np.uint64)
    drop = 23 - bits
    return ((((u + (1 << (drop - 1))) >> drop) << drop).astype(np.uint32)).view(np.float32)


def tf32(x):
    return round_mantissa(x, 10)


def bf16(x):
    return round_mantissa(x, 7)


def main():
    rng = np.random.default_rng(0)
    d = np.float64
    for K in (576, 4608, 25088):
        a = np.maximum(rng.standard_normal((256, K)) * np.abs(rng.standard_normal((256, K))), 0).astype(np.float32)
        b = (rng.standard_normal((K, 64)) * np.sqrt(2.0 / K)).astype(np.float32)
        exact = a.astype(d) @ b.astype(d)
        ah, bh = tf32(a), tf32(b)
        main_term = ah.astype(d) @ bh.astype(d)
        three = main_term + tf32(a - ah).astype(d) @ bh.astype(d) + ah.astype(d) @ tf32(b - bh).astype(d)
        mixed = main_term + bf16(a - ah).astype(d) @ bf16(bh).astype(d) + bf16(ah).astype(d) @ bf16(b - bh).astype(d)
        scale = np.abs(exact).max()
        err = lambda y: np.abs(y - exact).max() / scale
        print("K=%5d  3xTF32 %.2e   tf32+2xbf16 %.2e   1xTF32 %.2e   fp32 CPU matmul %.2e"
              % (K, err(three), err(mixed), err(main_term), err((a @ b).astype(d))))


if __name__ == "__main__":
    main()
