"""malva_b200 -- B200-native MALVA genotyping hot path (hand-written sm_100a kernels behind a C ABI)."""
from .api import (BF_ALT, BF_CONTEXT, KMAP_REF, GenotypeResult, KmerCounter, MalvaGpu, MalvaGpuError,  # noqa: F401
                  SignatureBatch, genotype_names, make_pool)
