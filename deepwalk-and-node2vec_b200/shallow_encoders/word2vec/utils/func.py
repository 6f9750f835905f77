"""
`pairwise_cosine_similarity` with the reference's signature (shallow_encoders/word2vec/utils/func.py:7-20: x / |x| times (y / |y|)^T),
computed on the device: row norms by `se_row_inv_norms`, the product by the tcgen05 tensor-core GEMM with the norms applied in its
epilogue (`se_cosine_similarity`, csrc/gemm.cu).  Inputs may live on the host, as in the reference's analysis tool, which passes
`model.input_embedding` (a CPU copy); they are moved to the current CUDA device for the computation and the result comes back where
`x` lives.  There is no CPU implementation: without a CUDA device the call raises.
"""
import torch

from shallow_encoders import _native as nat


def pairwise_cosine_similarity(x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
    """Similarity matrix [len(x), len(y)] of the rows of two 2-D tensors with the same number of columns."""
    assert x.dim() == 2 and y.dim() == 2 and x.shape[1] == y.shape[1], 'expected two matrices with the same number of columns'
    if not torch.cuda.is_available():
        raise RuntimeError('pairwise_cosine_similarity runs on the B200 kernels: no CUDA device, and there is no CPU fallback')
    dev = x.device if x.is_cuda else (y.device if y.is_cuda else torch.device('cuda', torch.cuda.current_device()))
    xd = x.detach().to(device=dev, dtype=torch.float32).contiguous()
    yd = y.detach().to(device=dev, dtype=torch.float32).contiguous()
    out = nat.cosine_similarity(xd, yd)
    return out if x.is_cuda else out.to(x.device)
