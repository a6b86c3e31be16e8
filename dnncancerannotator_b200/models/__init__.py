from . import tf_models  # noqa: F401  (same registry name as the reference: annotator/models/__init__.py:1)
