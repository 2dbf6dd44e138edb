"""Model zoo.  GPT-2 small/medium/large/XL (the only family the reference defines, example/model.py)."""
from .gpt2 import (GPTConfig, GPT2Model, standard_attention, flash_attention, CausalSelfAttention, MLP, Block,
                   PRESETS, gpt2_config)

__all__ = ["GPTConfig", "GPT2Model", "standard_attention", "flash_attention", "CausalSelfAttention", "MLP",
           "Block", "PRESETS", "gpt2_config"]
