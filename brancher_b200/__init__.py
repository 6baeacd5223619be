"""brancher_b200 -- B200-native drop-in for the Monte-Carlo ELBO / pathwise-gradient hot path of Brancher.

Same module / class / method names as the reference package `brancher` (variables, standard_variables,
functions, inference, gradient_estimators, optimizers, geometric_ranges, modules, config), so a script
written for the reference runs with `import brancher_b200.functions as BF`, etc.  The ELBO and its
gradient are evaluated ONLY by the hand-written sm_100a kernels behind `brancher_b200._cuda`
(include/brancher_cuda.h); there is no eager / CPU fallback for that path.
"""
__version__ = "0.1.0"
