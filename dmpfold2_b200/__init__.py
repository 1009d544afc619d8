"""dmpfold2_b200 -- B200-native (sm_100a) DMPfold2 inference engine; drop-in for `dmpfold`'s public API."""
from .predict import aln_to_coords, alns_to_coords, confidence_summary, run_dmpfold  # noqa: F401
