"""CPU restatement (TEST INFRASTRUCTURE -- only tests/, smoke() and bench.py's CPU legs may import it) of
  fft/src/fft.rs:6-60            dft / fft_evaluate / fft_interpolate (recursive radix-2, natural order)
  merkle_tree/src/merkle_tree.rs:31-214   MerkleTree over Keccak-256
in plain Python integers, function by function.

Third-party arithmetic the reference takes from crates that are not under /root/reference:
  * ark-ff 0.5.0 `FftField::get_root_of_unity(n)`: TWO_ADIC_ROOT_OF_UNITY = GENERATOR^((p - 1) / 2^TWO_ADICITY), squared
    TWO_ADICITY - log2(n) times.  ark-bn254 0.5.0 Fr: GENERATOR = 5, TWO_ADICITY = 28; Fq: 3, 1; ark-bls12-381 0.5.0 Fr: 7, 32
    (Cargo.lock pins the versions); for BN254 Fr this gives 19103219067921713944291392827692070036145651957329286315305642004821462161904,
    the TWO_ADIC_ROOT_OF_UNITY ark-bn254 publishes (tests/test_fft_merkle_oracle.py).  Beyond that constant PARITY IS UNPINNED for the choice of root: the reference's own tests
    (fft.rs:104-137) compare against powers of the same `get_root_of_unity`, i.e. they hold for any primitive root; what is
    pinned is that w has exact order n, that evaluate is the polynomial's value at w^j, and interpolate o evaluate = id.
  * sha3 0.10.8 Keccak256: oracle/pyref.py keccak256 (pinned by the public Keccak-256 KATs)."""
from typing import List, Sequence, Tuple

from . import pyref as R

GENERATOR = {R.BN254_FR: (5, 28), R.BN254_FQ: (3, 1), R.BLS12_381_FR: (7, 32)}


def get_root_of_unity(n: int, p: int) -> int:  # ark-ff FftField::get_root_of_unity
    g, s = GENERATOR[p]
    log_n = n.bit_length() - 1
    assert n == 1 << log_n and log_n <= s
    w = pow(g, (p - 1) >> s, p)
    for _ in range(s - log_n):
        w = w * w % p
    return w


def split_poly(poly: Sequence[int]) -> Tuple[List[int], List[int]]:  # fft.rs:62-68
    return list(poly[0::2]), list(poly[1::2])


def dft(values: Sequence[int], root: int, p: int) -> List[int]:  # fft.rs:6-29
    n = len(values)
    if n == 1:
        return list(values)
    even, odd = split_poly(values)
    root_sq = root * root % p
    y_even, y_odd = dft(even, root_sq, p), dft(odd, root_sq, p)
    y = [0] * n
    for j in range(n // 2):
        tw = pow(root, j, p)
        y[j] = (y_even[j] + tw * y_odd[j]) % p
        y[j + n // 2] = (y_even[j] - tw * y_odd[j]) % p
    return y


def fft_evaluate(coefficients: Sequence[int], p: int) -> List[int]:  # fft.rs:31-41
    n = len(coefficients)
    if n == 0 or n & (n - 1):
        raise ValueError("Length must be a power of 2")
    return dft(coefficients, get_root_of_unity(n, p), p)


def fft_interpolate(evaluations: Sequence[int], p: int) -> List[int]:  # fft.rs:43-60
    n = len(evaluations)
    if n == 0 or n & (n - 1):
        raise ValueError("Length must be a power of 2")
    omega_inv = pow(get_root_of_unity(n, p), p - 2, p)
    inv_n = pow(n, p - 2, p)
    return [c * inv_n % p for c in dft(evaluations, omega_inv, p)]


# ------------------------------------------------------------------ merkle_tree.rs
LEFT, RIGHT = 0, 1


def compute_hash(data: int, p: int) -> int:  # :201-206
    return int.from_bytes(R.keccak256(R.fq_vec_to_bytes([data % p])), "little") % p


def hash_pair(left: int, right: int, p: int) -> int:  # :208-214
    return int.from_bytes(R.keccak256(R.fq_vec_to_bytes([left % p]) + R.fq_vec_to_bytes([right % p])), "little") % p


class MerkleTree:
    def __init__(self, depth: int, p: int, inputs: Sequence[int] = ()):  # new :31-50 / new_with_inputs :52-84
        if len(inputs) > (1 << depth):
            raise ValueError("Too many inputs for tree depth")
        self.p, self.depth = p, depth
        self.leaves = [0] * (1 << depth)
        for i, x in enumerate(inputs):
            self.leaves[i] = compute_hash(x, p)
        self.tree: List[List[int]] = []
        cur = list(self.leaves)
        for _ in range(depth):
            cur = [hash_pair(cur[2 * i], cur[2 * i + 1], p) for i in range(len(cur) // 2)]
            self.tree.append(list(cur))

    def update_leaf(self, leaf_id: int, data: int, is_hash: bool) -> None:  # :86-132
        if leaf_id >= 1 << self.depth:
            raise ValueError("Invalid leaf ID")
        cur = data % self.p if is_hash else compute_hash(data, self.p)
        self.leaves[leaf_id] = cur
        index = leaf_id
        for level in range(self.depth):
            sib = self.leaves[index ^ 1] if level == 0 else self.tree[level - 1][index ^ 1]
            left, right = (cur, sib) if index % 2 == 0 else (sib, cur)
            cur = hash_pair(left, right, self.p)
            index //= 2
            self.tree[level][index] = cur

    def get_root_hash(self) -> int:  # :134-136
        return self.tree[self.depth - 1][0]

    def create_proof(self, data_to_prove: int, leaf_id: int) -> Tuple[int, List[Tuple[int, int]]]:  # :138-183
        if leaf_id >= 1 << self.depth:
            raise ValueError("Invalid leaf ID")
        if self.leaves[leaf_id] != compute_hash(data_to_prove, self.p):
            raise ValueError("Data does not match the leaf hash")
        proof, index = [], leaf_id
        for level in range(self.depth):
            sib = self.leaves[index ^ 1] if level == 0 else self.tree[level - 1][index ^ 1]
            proof.append((sib, RIGHT if index % 2 == 0 else LEFT))
            index //= 2
        return data_to_prove, proof

    def verify(self, proof: Tuple[int, List[Tuple[int, int]]]) -> bool:  # :185-199
        data, path = proof
        cur = compute_hash(data, self.p)
        for h, side in path:
            cur = hash_pair(h, cur, self.p) if side == LEFT else hash_pair(cur, h, self.p)
        return cur == self.get_root_hash()
