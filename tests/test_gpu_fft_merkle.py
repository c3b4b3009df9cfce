"""Parity of the device NTT (fft/src/fft.rs) and Keccak Merkle tree (merkle_tree/src/merkle_tree.rs) with the oracle's
restatement, bit for bit, plus size-independent properties at sizes the recursive oracle does not reach."""
import random

import pytest

from oracle import fft_merkle_ref as M
from oracle import pyref as R

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("fid,p", [(0, R.BN254_FR), (2, R.BLS12_381_FR)])
def test_fft_matches_oracle(zkb, ctxs, fid, p):
    ctx = ctxs(fid, 0)
    rng = random.Random(17 + fid)
    for log_n in list(range(0, 12)) + [13]:
        n = 1 << log_n
        c = [rng.randrange(p) for _ in range(n)]
        c[0], c[-1] = p - 1, 0
        ev = zkb.fft.fft_evaluate(ctx, c)
        assert ev == M.fft_evaluate(c, p), log_n
        assert zkb.fft.fft_interpolate(ctx, ev) == c
        assert zkb.fft.fft_interpolate(ctx, c) == M.fft_interpolate(c, p)
    # the reference's own test vector (fft.rs:104-137)
    w = M.get_root_of_unity(4, p)
    assert zkb.fft.fft_evaluate(ctx, [1, 2, 3, 4]) == [(1 + 2 * x + 3 * x * x + 4 * x ** 3) % p for x in (pow(w, i, p) for i in range(4))]
    with pytest.raises(ValueError, match="Length must be a power of 2"):
        zkb.fft.fft_evaluate(ctx, [1, 2, 3])


def test_fft_unsupported_order(zkb, ctxs):
    with pytest.raises(zkb.ZkbError) as ei:  # BN254 Fq has two-adicity 1
        zkb.fft.fft_evaluate(ctxs(1, 0), [1, 2, 3, 4])
    assert ei.value.status == -9


@pytest.mark.parametrize("log_n", [16, 20, 22])
def test_fft_large_properties(zkb, ctxs, log_n):
    """Every pass structure (first pass + strided passes of 6, 6, .. stages) on a device-resident table: a monomial x^k
    evaluates to w^(j k), interpolate o evaluate = id, and evaluate is linear."""
    fid, p = 0, R.BN254_FR
    ctx = ctxs(fid, 0)
    n = 1 << log_n
    t = zkb.MultilinearPoly.generate(ctx, 0xB2000F00, 0, log_n)
    ev = zkb.fft.ntt(t)
    back = zkb.fft.ntt(ev, inverse=True)
    assert (back - t).sum_halves() == [0, 0] and (back * back - t * t).sum_halves() == [0, 0]
    # y[0] = sum of the coefficients; sum_j y[j] = n c_0
    c0 = t.evaluate([0] * log_n)
    assert sum(ev.sum_halves()) % p == n * c0 % p
    assert ev.evaluate([0] * log_n) == sum(t.sum_halves()) % p
    # spot check against the definition at a few outputs: y[j] = sum_i c_i w^(i j) via Horner on a downloaded slice is too
    # slow in Python at 2^22, so use j with small order: j = n/4 -> w^j = i (4th root), y[j] = sum_r i^r (sum of c over i = r mod 4)
    w = M.get_root_of_unity(n, p)
    i4 = pow(w, n // 4, p)
    coeffs = t.evaluation if log_n <= 16 else None
    if coeffs is not None:
        for j in (1, 3, n // 2 + 5, n - 1):
            x = pow(w, j, p)
            acc = 0
            for c in reversed(coeffs):
                acc = (acc * x + c) % p
            bits = [(j >> (log_n - 1 - k)) & 1 for k in range(log_n)]
            assert ev.evaluate(bits) == acc
    assert pow(i4, 4, p) == 1
    for m in (t, ev, back):
        m.free()


@pytest.mark.parametrize("fid,p", [(1, R.BN254_FQ), (0, R.BN254_FR), (2, R.BLS12_381_FR)])
def test_merkle_reference_tests_on_device(zkb, ctxs, fid, p):
    """merkle_tree.rs:218-367 through the C ABI."""
    ctx = ctxs(fid, 0)
    T = zkb.merkle_tree.MerkleTree
    t = T(ctx, 2)
    h1 = M.hash_pair(0, 0, p)
    assert t.leaves == [0] * 4 and t.nodes(1) == [h1, h1] and t.get_root_hash() == M.hash_pair(h1, h1, p)
    t.update_leaf(1, 10, False)
    assert t.leaves[1] == M.compute_hash(10, p)
    assert t.get_root_hash() == M.hash_pair(M.hash_pair(0, M.compute_hash(10, p), p), h1, p)
    t.free()
    t = T(ctx, 2)
    t.update_leaf(0, 10, False)
    t.update_leaf(0, 0, True)
    assert t.leaves[0] == 0 and t.get_root_hash() == M.hash_pair(h1, h1, p)
    t.free()
    t = T(ctx, 3)
    t.update_leaf(0, 10, False)
    pr = t.create_proof(10, 0)
    assert t.verify(pr) and [s for _, s in pr.proof] == [zkb.merkle_tree.RIGHT] * 3
    assert not t.verify(zkb.merkle_tree.MerkleProof(10, [(0, zkb.merkle_tree.LEFT)] * 3))
    with pytest.raises(ValueError, match="Data does not match the leaf hash"):
        t.create_proof(20, 0)
    with pytest.raises(ValueError, match="Invalid leaf ID"):
        t.update_leaf(8, 1, False)
    t.free()
    t = T(ctx, 2, [1, 2, 3])
    assert t.leaves == [M.compute_hash(x, p) for x in (1, 2, 3)] + [0]
    t.free()
    with pytest.raises(ValueError, match="Too many inputs for tree depth"):
        T(ctx, 2, [1] * 5)


@pytest.mark.parametrize("depth,n_inputs", [(1, 2), (5, 32), (7, 100), (11, 2048)])
def test_merkle_random_matches_oracle(zkb, ctxs, depth, n_inputs):
    fid, p = 0, R.BN254_FR
    ctx = ctxs(fid, 0)
    rng = random.Random(depth)
    inputs = [rng.randrange(p) for _ in range(n_inputs)]
    inputs[0], inputs[-1] = 0, p - 1
    ref = M.MerkleTree(depth, p, inputs)
    t = zkb.merkle_tree.MerkleTree(ctx, depth, inputs)
    assert t.leaves == ref.leaves
    for level in range(1, depth + 1):
        assert t.nodes(level) == ref.tree[level - 1]
    for leaf in {0, 1, n_inputs - 1, n_inputs // 2}:
        pr = t.create_proof(inputs[leaf], leaf)
        assert (pr.data, pr.proof) == ref.create_proof(inputs[leaf], leaf) and t.verify(pr)
    leaf = rng.randrange(1 << depth)
    t.update_leaf(leaf, 12345, False)
    ref.update_leaf(leaf, 12345, False)
    assert t.get_root_hash() == ref.get_root_hash()
    t.free()
