from .training import EfficientDenseRetrievalTrainer  # noqa: F401
