"""Drop-in replacement for the reference's ``layers`` package (same class names, constructor and
``forward`` signatures, parameter names and shapes), backed by the sm_100a kernels."""
from .attention import BiDAFAttention, MultimodalAttentionDecoder, masked_softmax  # noqa: F401
from .encoding import Embedding, HighwayEncoder, ImageEmbedding, RNNEncoder  # noqa: F401
