"""Mirror of ``src/main/stat_result_vessel.py``: ``export_result(save_dir, test_config)`` for the vessel
pipeline (``pipeline_vessel.py`` imports it under this name)."""
from .stat_result import export_result_vessel as export_result  # noqa: F401
