from .ms_deform_attn_func import MSDeformAttnFunction, MSDeformAttnFunctionBF16  # noqa: F401
