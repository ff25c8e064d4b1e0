"""malva_b200 -- B200-native MALVA genotyping hot path (hand-written sm_100a kernels behind a C ABI)."""
from .api import (BF_ALT, BF_CONTEXT, KMAP_REF, GenotypeResult, MalvaGpu, MalvaGpuError, SignatureBatch,  # noqa: F401
                  genotype_names, make_pool)
