"""GPT-2 for the examples.  The classes live in ``tiny_deepspeed_b200.models.gpt2`` (same names, same parameter
registration order as the reference's example/model.py); this file keeps the reference's import path
``from example.model import GPTConfig, GPT2Model`` alive."""
from tiny_deepspeed_b200.models.gpt2 import (GPTConfig, GPT2Model, standard_attention, flash_attention,  # noqa: F401
                                             CausalSelfAttention, MLP, Block, PRESETS, gpt2_config)
