"""eyediseasesegmentation_b200 -- B200 (sm_100a) inference-and-scoring hot path of
duylebkHCM/EyeDiseaseSegmentation behind the reference's own Python interfaces.

Layout (mirrors the reference's ``src/main`` modules that sit on the path):
  archs/        ``MODEL_REGISTRY`` / ``get_model`` / ``get_preprocessing_fn`` (archs/__init__.py)
  tta.py        ``test_tta`` / ``tta_patches`` (src/main/tta.py)
  aucpr.py      ``get_auc`` / ``get_aucroc`` / ``plot_aucpr_curve`` / ``plot_aucroc_curve``
  util.py       ``make_grid`` / ``multigen`` / ``lesion_dict`` / ``save_output`` (util/base_utils.py)
  ttach_compat  ``aliases`` / ``SegmentationTTAWrapper`` (third-party ttach)
  csrc/         CUDA kernels + the C ABI declared in include/eds_b200.h
"""
__version__ = "0.1.0"
