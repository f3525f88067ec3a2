"""beast_tokenizer_b200 — B200-native (sm_100a) implementation of the BEAST tokenizer hot path,
behind the reference's Python API (Dont4rootMe/beast_tokenizer, package `beast`)."""
from .base_tokenizer import TokenizerBase
from .beast_bspline_tokenizer import BEASTBsplineTokenizer, CONFIG_FILENAME
from .beast_bspline_bpe_tokenizer import BEASTBsplineBPETokenizer
from .beast_bpe_trainer import FIGBPE, FIGBPEState
from .bpe_model import B200ByteLevelBPE
from ._lib import BeastB200Error

__all__ = ["TokenizerBase", "BEASTBsplineTokenizer", "BEASTBsplineBPETokenizer", "FIGBPE", "FIGBPEState",
           "B200ByteLevelBPE", "CONFIG_FILENAME", "BeastB200Error"]
