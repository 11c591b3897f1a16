"""Import shim (golden-vector generation only): routes ``x_transformers`` to the oracle restatement."""
from oracle.xt_encoder import Decoder, Encoder  # noqa: F401
