"""B200-native GP log-marginal-likelihood hot path of AutoGP.jl (Gram build -> Cholesky ->
solve -> logdet), behind the reference's own operator names.  CUDA only; see DESIGN.md."""
from . import _lib  # noqa: F401
from .gp import (  # noqa: F401
    BinaryOpNode, ChangePoint, Constant, Engine, GammaExponential, LeafNode, Linear, Node, Periodic, Plus,
    SquaredExponential, Times, WhiteNoise, compute_cov_matrix, compute_cov_matrix_vectorized, default_engine,
    depth, encode_program, eval_cov, infer_gp_sum, marginal_quantiles, predictive_mvn, size, split_kernel_sop, unroll,
)
from .model import (  # noqa: F401
    JITTER, PosDefException, log_marginal_likelihood_grads, log_marginal_likelihoods, log_marginal_likelihoods_info,
    mvnormal_logpdf, predictive_logpdfs, transform_param, transform_param_grad, untransform_param,
)
from . import rejuvenate, smc, tree_moves  # noqa: F401

__version__ = "0.1.0"
