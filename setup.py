"""Packaging of the drop-in: mirrors what the reference's setup.py provides (a package, the `dmpfold` console script,
numpy as the only hard dependency besides torch) -- plus the one thing the reference does not need, the native build.

`pip install .` compiles dmpfold2_b200/csrc/*.cu for sm_100a with nvcc (python -m dmpfold2_b200.build) and ships the
resulting libdmp2.so inside the package.  For development and for the GPU tests the library is built IN-TREE with the
same command and loaded from the source checkout; nothing in this repository depends on an installed copy.
"""
import os
import sys

import setuptools
from setuptools.command.build_py import build_py

HERE = os.path.dirname(os.path.abspath(__file__))


class BuildWithCuda(build_py):
    def run(self):
        sys.path.insert(0, HERE)
        from dmpfold2_b200.build import build          # nvcc -gencode arch=compute_100a,code=sm_100a -> libdmp2.so
        build()
        super().run()


with open(os.path.join(HERE, 'README.md')) as fh:
    long_description = fh.read()

setuptools.setup(
    name='dmpfold2-b200',
    version='0.1',
    description='B200-native (sm_100a) inference engine for the DMPfold2 protein structure predictor; drop-in for the '
                'dmpfold CLI and dmpfold.aln_to_coords()',
    long_description=long_description,
    long_description_content_type='text/markdown',
    packages=['dmpfold2_b200', 'dmpfold'],             # `dmpfold` = import-name alias of the reference package
    package_data={'dmpfold2_b200': ['libdmp2.so', 'csrc/*.cu', 'csrc/*.cuh', 'trained_model/*.pt']},
    scripts=['bin/dmpfold'],
    install_requires=['numpy', 'torch'],
    python_requires='>=3.9',
    cmdclass={'build_py': BuildWithCuda},
    classifiers=['Programming Language :: Python :: 3', 'Environment :: GPU :: NVIDIA CUDA',
                 'Topic :: Scientific/Engineering :: Bio-Informatics'],
)
