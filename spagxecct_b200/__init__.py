"""spagxecct_b200 -- B200-native (sm_100a) Step-2 per-SNP testing loop of SPAGxECCT.

The package is a thin host mirror of the reference's exported R functions over the C ABI of
libspagxe_cuda.so (include/spagxe.h).  There is no CPU fallback.
"""
from .api import (COLS_CCT, COLS_LOCAL, COLS_MIX, COLS_PLUS, RStop, Result, SPAGxE_CCT, SPAGxE_CCT_one_SNP, SPAGxE_Plus,  # noqa: F401
                  SPAGxE_Plus_one_SNP, SPAGxEmix_CCT, SPAGxEmix_CCT_one_SNP, SPAGxEmixCCT_localance,
                  SPAGxEmixCCT_localance_one_SNP, check_input_Resid, CCT, COLS_MIX_PLUS, SPAGxEmix_CCT_Plus,
                  SPAGxEmix_CCT_Plus_one_SNP)
from ._lib import Context, MultiContext, SpagxeError, make_params  # noqa: F401

__version__ = "0.1.0"
