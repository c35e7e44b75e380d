"""pathmatfac.jl_b200 -- B200-native fit-loop hot path of PathMatFac.jl.

Holds the CUDA kernels + C ABI (``csrc/``, built into ``libpmf.so``) and the host-side
mirror of the reference interface for that path (``PathMatFacModel``, ``mf_fit``,
``mf_fit_adapt_lr``).  Import it as ``pathmatfac_b200`` (the directory name carries a dot;
``pathmatfac_b200.py`` at the repo root is the import alias)."""
from .fit import (AdaGrad, Engine, compute_M_estimates, cpu, gpu, init_logsigma, init_mu, mf_fit, mf_fit_adapt_lr,
                  reweight_col_losses, theta_delta_em)
from .layers import (BatchArray, BatchScale, BatchShift, ColScale, ColShift, FrozenLayer,
                     ViewableComposition, construct_model_layers, freeze_layer, unfreeze_layer)
from .model import CompositeNoise, MatFacModel, PathMatFacModel
from .transform import transform
from .model_io import (load_batches, load_model, load_omic_data, model_arrays, read_model_hdf, save_model, save_omic_data,
                       save_transformed, write_model_arrays, write_model_to_hdf)
from . import h5lite
from .prep_pathways import prep_pathway_featuresets, prep_pathway_graphs
from . import staging      # fit!, basic_fit!, init_batch_effects!, ... (kept in their namespace: `fit` is also a submodule)
from .postfit import (init_ordinal_thresholds, reorder_by_importance, reorder_reg, reweight_eb, rotate_by_svd,
                      whiten)
from .regularizers import (ARDRegularizer, BatchArrayReg, ColParamReg, CompositeRegularizer,
                           FeatureSetARDReg, FrozenRegularizer, GroupRegularizer, L2Regularizer,
                           NetworkRegularizer, SelectiveL1Reg, SequenceReg, ZeroReg,
                           construct_featureset_ard, construct_layer_reg, construct_X_reg,
                           construct_Y_reg, freeze_reg, unfreeze_reg)

__all__ = [n for n in dir() if not n.startswith("_")]
