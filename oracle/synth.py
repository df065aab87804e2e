"""Back-compat shim: the seeded input generator lives in the package (hspose_b200.synth);
it is a data generator, not part of the oracle."""
from hspose_b200.synth import *  # noqa: F401,F403
from hspose_b200.synth import fill_params, synth_batch  # noqa: F401
