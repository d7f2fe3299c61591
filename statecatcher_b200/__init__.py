"""statecatcher_b200 — B200-native LucyRNN + CTC (+RNN-T) training hot path.

Python host code over hand-written sm_100a CUDA kernels behind a C-ABI
(include/statecatcher_b200.h).  No Triton, no multi-backend dispatch, no CPU fallback.
"""
from .lucyrnn_conf import LucyRNNConfig
from .lucyrnn import LucyRNN, LucyRNNCell, LucyRNNtriton
from .ctc import CTCLoss, ctc_loss, ctc_loss_from_logits
from .rnnt import RNNTCompactPredictorJoiner, RNNTFusedHead, RNNTLoss, RNNTPredictorJoiner, rnnt_loss
from .decoder import ctc_greedy_decoder
from .frontend import MFCC, MelDB, compute_frame_mask, featurize, frame_mask_and_lens, make_frontend
from .glue import GraphedStreamingEncoder, GraphedTrainStep, LucyASRModel, SegmentPrefetcher, assert_all_detached, compute_loss, detach_states

__all__ = ["LucyRNNConfig", "LucyRNN", "LucyRNNCell", "LucyRNNtriton", "CTCLoss", "ctc_loss",
           "ctc_loss_from_logits", "LucyASRModel", "compute_loss", "detach_states", "SegmentPrefetcher", "GraphedStreamingEncoder", "GraphedTrainStep",
           "assert_all_detached", "RNNTLoss", "RNNTPredictorJoiner", "RNNTCompactPredictorJoiner", "RNNTFusedHead", "rnnt_loss", "ctc_greedy_decoder",
           "MFCC", "MelDB", "make_frontend", "compute_frame_mask", "frame_mask_and_lens", "featurize"]
