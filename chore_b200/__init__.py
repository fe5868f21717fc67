"""chore_b200: the CHORE hot path (hourglass encoder, pixel-aligned point query, SMPL-H LBS +
rigid object fitting step) as hand-written sm_100a CUDA kernels behind the reference's own
Python interfaces.  See DESIGN.md; the C ABI is include/chore_b200.h."""
from ._lib import (HEAD_ALL, HEAD_CENTERS, HEAD_DF, HEAD_PARTS, HEAD_PCA, ChoreError, Handle, get_handle,
                   launch_count, load_library)
from .fitter import FusedAdam, FusedFitSteps, GraphedStep, ReconFitterBase, ReconFitterBehave, backward_to
from .data import TestData
from .generator import Generator
from .net import CHORE
from .smpl import SMPLHLayer, SMPLPyTorchWrapperBatch, SMPLPyTorchWrapperBatchSplitParams

__all__ = ["CHORE", "Generator", "ReconFitterBase", "ReconFitterBehave", "GraphedStep", "FusedFitSteps", "FusedAdam", "TestData", "backward_to", "SMPLHLayer", "SMPLPyTorchWrapperBatch",
           "SMPLPyTorchWrapperBatchSplitParams", "Handle", "get_handle", "load_library", "launch_count", "ChoreError",
           "HEAD_ALL", "HEAD_DF", "HEAD_PCA", "HEAD_PARTS", "HEAD_CENTERS"]
