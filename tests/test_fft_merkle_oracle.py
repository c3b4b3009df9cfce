"""The oracle's restatement of fft/src/fft.rs and merkle_tree/src/merkle_tree.rs against the reference's own unit tests
(fft.rs:88-137, merkle_tree.rs:218-367) and the published ark-bn254 root of unity.  CPU only."""
import pytest

from oracle import fft_merkle_ref as M
from oracle import pyref as R

P, Q = R.BN254_FR, R.BN254_FQ


def test_two_adic_root_is_the_published_constant():
    assert pow(5, (P - 1) >> 28, P) == 19103219067921713944291392827692070036145651957329286315305642004821462161904
    for n in (1, 2, 4, 1 << 10, 1 << 28):
        w = M.get_root_of_unity(n, P)
        assert pow(w, n, P) == 1 and (n == 1 or pow(w, n // 2, P) == P - 1)


def test_splits_poly_correctly():  # fft.rs:88-102
    assert M.split_poly([2, P - 14, 2, 1]) == ([2, 2], [P - 14, 1])


def test_evaluates_and_interpolates_poly():  # fft.rs:104-137
    coeffs = [1, 2, 3, 4]
    ev = M.fft_evaluate(coeffs, P)
    w = M.get_root_of_unity(4, P)
    assert ev == [(1 + 2 * x + 3 * x * x + 4 * x ** 3) % P for x in (pow(w, i, P) for i in range(4))]
    assert M.fft_interpolate(ev, P) == coeffs
    with pytest.raises(ValueError, match="Length must be a power of 2"):
        M.fft_evaluate([1, 2, 3], P)


def test_merkle_reference_tests():  # merkle_tree.rs:218-367
    t = M.MerkleTree(2, Q)
    assert len(t.leaves) == 4 and [len(l) for l in t.tree] == [2, 1] and t.leaves == [0] * 4
    h1 = M.hash_pair(0, 0, Q)
    assert t.get_root_hash() == M.hash_pair(h1, h1, Q)
    t.update_leaf(1, 10, False)  # test_update_leaf
    assert t.leaves[1] == M.compute_hash(10, Q)
    assert t.get_root_hash() == M.hash_pair(M.hash_pair(0, t.leaves[1], Q), M.hash_pair(0, 0, Q), Q)
    t = M.MerkleTree(2, Q)  # test_delete_leaf
    t.update_leaf(0, 10, False)
    t.update_leaf(0, 0, True)
    assert t.leaves[0] == 0 and t.get_root_hash() == M.hash_pair(h1, h1, Q)
    t = M.MerkleTree(3, Q)  # test_proof_and_verify
    t.update_leaf(0, 10, False)
    assert t.verify(t.create_proof(10, 0))
    t = M.MerkleTree(2, Q)  # test_verify_invalid_proof
    assert not t.verify((10, [(0, M.LEFT)] * 2))
    t.update_leaf(0, 10, False)  # test_create_proof_invalid_data
    with pytest.raises(ValueError, match="Data does not match the leaf hash"):
        t.create_proof(20, 0)
    t = M.MerkleTree(2, Q, [1, 2, 3])  # test_new_with_inputs
    assert t.leaves[:3] == [M.compute_hash(x, Q) for x in (1, 2, 3)] and t.leaves[3] == 0
    with pytest.raises(ValueError, match="Too many inputs for tree depth"):
        M.MerkleTree(2, Q, [1] * 5)
