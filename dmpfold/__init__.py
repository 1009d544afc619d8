"""`dmpfold` -- import-name alias of the B200-native engine, so that code written against the reference package
(`from dmpfold import aln_to_coords`, reference dmpfold/__init__.py:1) runs unchanged on dmpfold2_b200."""
from dmpfold2_b200.predict import aln_to_coords, run_dmpfold  # noqa: F401
