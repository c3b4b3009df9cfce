"""zkb200 -- B200-native sumcheck / GKR prover engine behind the crate API of
obah/zk-research-implementations.  The directory name carries a hyphen (it is
the reference's name), so import it with
    importlib.import_module("zk-research-implementations_b200")
Modules mirror the reference crates: multilinear_polynomial, sum_check_protocol,
gkr_circuit, gkr_protocol, fiat_shamir, univariate_polynomial, kzg (pcs), fft, merkle_tree; `engine` is the
ctypes binding of libzkb200.so (include/zkb200.h).  No CPU fallback exists."""
from . import engine  # noqa: F401
from . import fft, fiat_shamir, gkr_circuit, gkr_protocol, kzg, merkle_tree, multilinear_polynomial, sum_check_protocol, univariate_polynomial  # noqa: F401
from .engine import BLS12_381_FR, BN254_FQ, BN254_FR, MODE_COMPAT, MODE_FULL, Context, ZkbError  # noqa: F401
from .multilinear_polynomial import MultilinearPoly, Operation, ProductPoly, SumPoly  # noqa: F401
