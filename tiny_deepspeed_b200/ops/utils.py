"""Dtype helpers of the op layer (reference `tiny_deepspeed/core/module/ops/utils.py:10-16` maps torch dtypes to Triton
dtypes and lists accumulator dtypes; Triton is gone, the accumulator table stays)."""
import torch

# input dtype -> dtype the kernels accumulate in
supported_acc_dtypes = {
    torch.float16: torch.float32,
    torch.bfloat16: torch.float32,
    torch.float32: torch.float32,
    torch.int8: torch.int32,
}

# dtypes the sm_100a kernels accept for activations / parameters (bf16 -> kind::f16 GEMMs, fp32 -> kind::tf32 GEMMs on the fp32 tensors in place)
kernel_dtypes = (torch.bfloat16, torch.float32)


def acc_dtype(dtype: torch.dtype) -> torch.dtype:
    try:
        return supported_acc_dtypes[dtype]
    except KeyError:
        raise TypeError(f"unsupported dtype {dtype}") from None


def to_kernel_dtype_code(dtype: torch.dtype) -> int:
    """Code used across the C++ boundary (csrc/kernels.h: kBF16 = 0, kF32 = 1)."""
    if dtype == torch.bfloat16:
        return 0
    if dtype == torch.float32:
        return 1
    raise TypeError(f"unsupported dtype {dtype}")
