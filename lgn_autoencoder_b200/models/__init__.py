"""Models (API of the reference's lgn/models/__init__.py:1-5)."""
from .lgn_cg import LGNCG
from .lgn_decoder import LGNDecoder
from .lgn_encoder import LGNEncoder
from .lgn_levels import CGMLP, LGNNodeLevel

__all__ = ["LGNEncoder", "LGNDecoder", "LGNCG", "LGNNodeLevel", "CGMLP"]
