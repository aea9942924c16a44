"""gam_ngs_b200 - B200-native (sm_100a) implementation of GAM-NGS's contig-pair alignment
hot path: BandedSmithWaterman::find_alignment
(/root/reference/lib/src/alignment/banded_smith_waterman.cc:69-323) behind the reference's
own call interface.  See DESIGN.md and INTEGRATION.md."""
import os as _os

# A context drives ~20 CUDA streams per device; the default of 8 hardware queues makes some of them share a
# queue (false dependencies between a chunk and the contig upload, see gamx_create).  Read when the CUDA
# context is created, so it has to be in the environment before the first CUDA call of the process.
_os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

from .capi import (Context, GamxError, Plan, make_jobs, make_hits_jobs, load_library, JOB_DTYPE, RESULT_DTYPE,  # noqa: F401
                   MODE_SCORE, MODE_ENDPOINTS, MODE_FULL, JOB_OK, JOB_EMPTY, JOB_OUT_OF_RANGE,
                   JOB_UNDEFINED, DEFAULT_BAND, DEFAULT_GAP)
from .aligner import BandedSmithWaterman, Contig, MyAlignment  # noqa: F401
