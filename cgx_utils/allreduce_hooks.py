from torch_cgx_b200.parallel.hooks import (  # noqa: F401
    COMPRESSION_BUCKET_SIZE,
    COMPRESSION_MINIMAL_SIZE,
    COMPRESSION_QUANTIZATION_BITS,
    VALUE_NO_COMPRESS,
    CGXState,
    _allreduce_fut,
    cgx_hook,
)
