from .runtime_tuner import RuntimeAutoTuner

__all__ = ["RuntimeAutoTuner"]
