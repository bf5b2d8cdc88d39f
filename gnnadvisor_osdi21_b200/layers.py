"""Autograd layer over the extension surface -- host-side mirror of GNNAdvisor/gnn_conv.py.

Same class names and call conventions as the reference (ScatterAndGather :7-28, GNNAFunction
:31-78, GCNConv :80-98, GNNAFunction_GIN :101-126, GINConv :128-147) so a script written against
the reference's gnn_conv can import this module instead.  `inputInfo` is any object carrying
row_pointers, column_index, degrees, partPtr, part2Node, partSize, dimWorker, warpPerBlock
(param.InputProperty here, inputProperty in the reference).
"""
import math

import torch

from . import ops as GNNA


def _graph(info):
    return (info.row_pointers, info.column_index)


def _tune(info):
    return (info.partSize, info.dimWorker, info.warpPerBlock)


class ScatterAndGather(torch.autograd.Function):
    """out = A @ X; the graph is undirected so the backward is the same aggregation (gnn_conv.py:7-28)."""

    @staticmethod
    def forward(ctx, X, inputInfo):
        ctx.inputInfo = inputInfo
        ctx.tune = _tune(inputInfo)
        return GNNA.SAG(X, *_graph(inputInfo), inputInfo.degrees, inputInfo.partPtr, inputInfo.part2Node, *ctx.tune)

    @staticmethod
    def backward(ctx, d_output):
        info = ctx.inputInfo
        d_input = GNNA.SAG(d_output.contiguous(), *_graph(info), info.degrees, info.partPtr, info.part2Node, *ctx.tune)
        return d_input, None


class GNNAFunction(torch.autograd.Function):
    """GCN layer: update then aggregate (gnn_conv.py:31-78)."""

    @staticmethod
    def forward(ctx, X, weight, inputInfo):
        ctx.save_for_backward(X, weight)
        ctx.inputInfo = inputInfo
        ctx.tune = _tune(inputInfo)
        return GNNA.forward(X, weight, *_graph(inputInfo), inputInfo.degrees,
                            inputInfo.partPtr, inputInfo.part2Node, *ctx.tune)[0]

    @staticmethod
    def backward(ctx, d_output):
        X, weight = ctx.saved_tensors
        info = ctx.inputInfo
        # unlike the reference (kernel.cu:472 runs G @ W^T unconditionally) the input gradient is only
        # computed when autograd asks for it: the first layer's features need none
        d_input, d_weight = GNNA.backward(d_output.contiguous(), X, weight, *_graph(info), info.degrees,
                                          info.partPtr, info.part2Node, *ctx.tune,
                                          need_d_input=ctx.needs_input_grad[0])
        return d_input, d_weight, None


class GNNAFunctionMixed(torch.autograd.Function):
    """GNNAFunction with the gathered matrix stored as bf16 (extension: BASELINE.json's "Reddit GCN D=64 bf16"
    configuration; dense products, accumulation, outputs and gradients stay fp32)."""

    @staticmethod
    def forward(ctx, X, weight, inputInfo):
        ctx.save_for_backward(X, weight)
        ctx.inputInfo = inputInfo
        ctx.tune = _tune(inputInfo)
        return GNNA.forward_mixed(X, weight, *_graph(inputInfo), inputInfo.degrees,
                                  inputInfo.partPtr, inputInfo.part2Node, *ctx.tune)[0]

    @staticmethod
    def backward(ctx, d_output):
        X, weight = ctx.saved_tensors
        info = ctx.inputInfo
        d_input, d_weight = GNNA.backward_mixed(d_output.contiguous(), X, weight, *_graph(info), info.degrees,
                                                info.partPtr, info.part2Node, *ctx.tune,
                                                need_d_input=ctx.needs_input_grad[0])
        return d_input, d_weight, None


class GNNAFunction_GIN(torch.autograd.Function):
    """GIN layer: aggregate then update; the aggregated features are what backward needs (gnn_conv.py:101-126)."""

    @staticmethod
    def forward(ctx, X, weight, inputInfo, eplison):
        X_prime, X_agg = GNNA.forward_gin(X, weight, *_graph(inputInfo), eplison,
                                          inputInfo.partPtr, inputInfo.part2Node, *_tune(inputInfo))
        ctx.save_for_backward(X_agg, weight)
        ctx.inputInfo = inputInfo
        ctx.tune = _tune(inputInfo)
        ctx.eplison = eplison
        return X_prime

    @staticmethod
    def backward(ctx, d_output):
        X_agg, weight = ctx.saved_tensors
        info = ctx.inputInfo
        d_input, d_weight = GNNA.backward_gin(d_output.contiguous(), X_agg, weight, *_graph(info), ctx.eplison,
                                              info.partPtr, info.part2Node, *ctx.tune,
                                              need_d_input=ctx.needs_input_grad[0])
        return d_input, d_weight, None, None


class GNNAFunction_GIN_Mixed(torch.autograd.Function):
    """GNNAFunction_GIN with bf16 gathered rows (extension; accumulation, saved X_agg, products and gradients fp32)."""

    @staticmethod
    def forward(ctx, X, weight, inputInfo, eplison):
        X_prime, X_agg = GNNA.forward_gin_mixed(X, weight, *_graph(inputInfo), eplison,
                                                inputInfo.partPtr, inputInfo.part2Node, *_tune(inputInfo))
        ctx.save_for_backward(X_agg, weight)
        ctx.inputInfo = inputInfo
        ctx.tune = _tune(inputInfo)
        ctx.eplison = eplison
        return X_prime

    @staticmethod
    def backward(ctx, d_output):
        X_agg, weight = ctx.saved_tensors
        info = ctx.inputInfo
        d_input, d_weight = GNNA.backward_gin_mixed(d_output.contiguous(), X_agg, weight, *_graph(info), ctx.eplison,
                                                    info.partPtr, info.part2Node, *ctx.tune,
                                                    need_d_input=ctx.needs_input_grad[0])
        return d_input, d_weight, None, None


def fused_tile_supported(width, gather_dtype):
    """Widths of the AGGREGATED matrix the fused aggregate -> tcgen05 tile exists for (csrc/fused_gemm.cu)."""
    return width in ((64, 128, 256) if gather_dtype == "bf16" else (32, 64, 128))


class GNNAFunction_GIN_Fused(torch.autograd.Function):
    """GNNAFunction_GIN whose forward runs aggregation and X_agg @ W in ONE kernel: the gathered 128-row tile stays in
    shared memory and feeds a bf16 tcgen05.mma with the accumulator in TMEM (extension; BASELINE.json's "fused
    aggregate+X*W tensor-core tile").  X_agg (saved for backward, gnn_conv.py:109) is still written, in exact fp32; the
    product has bf16-rounded operands.  gather_bf16: the gathered rows travel as bf16 too.  Backward is the unfused one
    (there the product comes BEFORE the aggregation, kernel.cu:710-738)."""

    @staticmethod
    def forward(ctx, X, weight, inputInfo, eplison, gather_bf16):
        Xg = GNNA.scale_rows_bf16(X) if gather_bf16 else X
        X_prime, X_agg = GNNA.forward_gin_fused(Xg, weight, *_graph(inputInfo), eplison,
                                                inputInfo.partPtr, inputInfo.part2Node, *_tune(inputInfo))
        ctx.save_for_backward(X_agg, weight)
        ctx.inputInfo, ctx.tune, ctx.eplison, ctx.gather_bf16 = inputInfo, _tune(inputInfo), eplison, gather_bf16
        return X_prime

    @staticmethod
    def backward(ctx, d_output):
        X_agg, weight = ctx.saved_tensors
        info = ctx.inputInfo
        fn = GNNA.backward_gin_mixed if ctx.gather_bf16 else GNNA.backward_gin
        d_input, d_weight = fn(d_output.contiguous(), X_agg, weight, *_graph(info), ctx.eplison,
                               info.partPtr, info.part2Node, *ctx.tune, need_d_input=ctx.needs_input_grad[0])
        return d_input, d_weight, None, None, None


class GNNAFunction_FusedBackward(torch.autograd.Function):
    """GNNAFunction (GCN) whose BACKWARD fuses aggregate -> product: G = Ahat @ dOut and dX = G @ W^T in one kernel
    (kernel.cu:436-472 is exactly aggregation followed by torch::mm; SURVEY.md F2: for GCN the product after the
    aggregation sits in the backward pass).  G is still written (dW = X^T @ G needs it, :473).  Forward as the unfused layer."""

    @staticmethod
    def forward(ctx, X, weight, inputInfo, gather_bf16):
        ctx.save_for_backward(X, weight)
        ctx.inputInfo, ctx.tune, ctx.gather_bf16 = inputInfo, _tune(inputInfo), gather_bf16
        fn = GNNA.forward_mixed if gather_bf16 else GNNA.forward
        return fn(X, weight, *_graph(inputInfo), inputInfo.degrees, inputInfo.partPtr, inputInfo.part2Node, *ctx.tune)[0]

    @staticmethod
    def backward(ctx, d_output):
        X, weight = ctx.saved_tensors
        info = ctx.inputInfo
        d_output = d_output.contiguous()
        if not ctx.needs_input_grad[0]:          # first layer: nothing to fuse, only G and dW
            fn = GNNA.backward_mixed if ctx.gather_bf16 else GNNA.backward
            _, d_weight = fn(d_output, X, weight, *_graph(info), info.degrees, info.partPtr, info.part2Node, *ctx.tune,
                             need_d_input=False)
            return None, d_weight, None, None
        d_input, d_weight = GNNA.backward_fused(d_output, X, weight, *_graph(info), info.degrees, info.partPtr, info.part2Node,
                                                *ctx.tune, gather_bf16=ctx.gather_bf16)
        return d_input, d_weight, None, None


class _ConvBase(torch.nn.Module):
    def __init__(self, input_dim, output_dim, gather_dtype="fp32", fused=False):
        """gather_dtype="bf16" (extension, not in the reference): neighbour rows travel as bf16, everything else fp32.
        fused=True (extension): where the layer computes aggregate -> dense product (GIN forward, GCN backward) and the
        aggregated width has a fused tile (fused_tile_supported), both run in ONE kernel with the product on the tensor
        cores (bf16 operands, fp32 accumulate); other shapes silently keep the unfused operators."""
        super().__init__()
        if gather_dtype not in ("fp32", "bf16"):
            raise ValueError("gather_dtype must be 'fp32' or 'bf16'")
        self.gather_dtype = gather_dtype
        self.fused = bool(fused)
        # randn first, like the reference (gnn_conv.py:83, 131): the same torch seed then gives the same initial weights
        self.weights = torch.nn.Parameter(torch.randn(input_dim, output_dim))
        self.reset_parameters()

    def reset_parameters(self):
        # U(-1/sqrt(out), 1/sqrt(out))  (gnn_conv.py:86-88, 136-138)
        bound = 1.0 / math.sqrt(self.weights.size(1))
        with torch.no_grad():
            self.weights.uniform_(-bound, bound)


class GCNConv(_ConvBase):
    def forward(self, X, inputInfo):
        # the aggregated matrix of the backward pass is d_output: [N, output_dim]
        if self.fused and X.requires_grad and fused_tile_supported(self.weights.shape[1], self.gather_dtype) \
                and self.weights.shape[0] <= 256:
            return GNNAFunction_FusedBackward.apply(X, self.weights, inputInfo, self.gather_dtype == "bf16")
        fn = GNNAFunctionMixed if self.gather_dtype == "bf16" else GNNAFunction
        return fn.apply(X, self.weights, inputInfo)


class GINConv(_ConvBase):
    def __init__(self, input_dim, output_dim, gather_dtype="fp32", fused=False):
        super().__init__(input_dim, output_dim, gather_dtype, fused)
        self.eplison = 0.5          # fixed in the reference (gnn_conv.py:132); spelling kept

    def forward(self, X, inputInfo):
        # the aggregated matrix of the forward pass is X: [N, input_dim]
        if self.fused and fused_tile_supported(self.weights.shape[0], self.gather_dtype) and self.weights.shape[1] <= 256:
            return GNNAFunction_GIN_Fused.apply(X, self.weights, inputInfo, self.eplison, self.gather_dtype == "bf16")
        fn = GNNAFunction_GIN_Mixed if self.gather_dtype == "bf16" else GNNAFunction_GIN
        return fn.apply(X, self.weights, inputInfo, self.eplison)
