"""The proof wire format (zkb_proof_encode / zkb_proof_decode, host only): round trips, an independent Python
restatement of the layout, a pinned golden encoding, and rejection of malformed or non-canonical bytes."""
import random
import struct

import pytest

from oracle import pyref as R

FIELDS = [(0, R.BN254_FR), (1, R.BN254_FQ), (2, R.BLS12_381_FR)]


def restate(field, kind, msgs, claimed):
    out = b"ZKBP" + bytes([1, field, kind, 0]) + struct.pack("<I", len(msgs)) + claimed.to_bytes(32, "little")
    for m in msgs:
        out += bytes([len(m)]) + b"".join(x.to_bytes(32, "little") for x in m)  # fq_vec_to_bytes per element
    return out


@pytest.mark.parametrize("fid,p", FIELDS)
def test_round_trip_and_layout(zkb, fid, p):
    S = zkb.sum_check_protocol
    rng = random.Random(fid)
    # GkrProof: trimmed coefficient vectors of every length, including empty
    msgs = [[rng.randrange(p) for _ in range(l)] for l in (3, 0, 2, 4, 1, 3)]
    gp = S.GkrProof([zkb.univariate_polynomial.UnivariatePoly(m, fid) for m in msgs], rng.randrange(p), [])
    b = S.proof_to_bytes(gp, fid)
    assert b == restate(fid, 2, msgs, gp.claimed_sum)
    back, f2 = S.proof_from_bytes(b)
    assert f2 == fid and back.claimed_sum == gp.claimed_sum and [q.coefficients for q in back.proof_polynomials] == msgs
    # Proof of the plain sumcheck
    pm = [[rng.randrange(p), rng.randrange(p)] for _ in range(5)]
    pp = S.Proof(pm, rng.randrange(p))
    b1 = S.proof_to_bytes(pp, fid)
    assert b1 == restate(fid, 1, pm, pp.claimed_sum)
    back1, _ = S.proof_from_bytes(b1)
    assert back1.proof_polynomials == pm and back1.claimed_sum == pp.claimed_sum
    # empty proof
    assert S.proof_from_bytes(S.proof_to_bytes(S.Proof([], 7), fid))[0].proof_polynomials == []


def test_golden_bytes(zkb):
    """The reference's own composed-sumcheck vector (sum_check_protocol.rs:225-245): round polynomial [20, 28, 20]."""
    S = zkb.sum_check_protocol
    gp = S.GkrProof([zkb.univariate_polynomial.UnivariatePoly([20, 28, 20], 1)], 88, [])
    want = ("5a4b4250" "01" "01" "02" "00" "01000000" + (88).to_bytes(32, "little").hex() + "03"
            + (20).to_bytes(32, "little").hex() + (28).to_bytes(32, "little").hex() + (20).to_bytes(32, "little").hex())
    assert S.proof_to_bytes(gp, 1).hex() == want


def test_malformed_bytes_are_rejected(zkb):
    S = zkb.sum_check_protocol
    p = R.BN254_FR
    good = S.proof_to_bytes(S.Proof([[1, 2], [3, 4]], 3), 0)
    bad = [good[:-1], good + b"\0", b"XKBP" + good[4:], good[:4] + b"\2" + good[5:], good[:6] + b"\7" + good[7:],
           good[:12] + p.to_bytes(32, "little") + good[44:],                # claimed_sum == p: not canonical
           good[:45] + (p + 5).to_bytes(32, "little") + good[77:],          # a message element >= p
           good[:44] + b"\3" + good[45:]]                                   # kind 1 with a 3-element message
    for b in bad:
        with pytest.raises(zkb.ZkbError):
            S.proof_from_bytes(b)
    assert S.proof_from_bytes(good)[0].proof_polynomials == [[1, 2], [3, 4]]
