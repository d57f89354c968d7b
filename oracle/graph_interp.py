"""fp32 torch interpreter of the product's layer graphs (TEST INFRASTRUCTURE).

Executes gdn_pytorch_b200.graph.Graph unit by unit with torch.nn.functional, returning EVERY named tensor, so a
GPU mismatch can be traced to the first unit that diverges.  The graph description itself is validated against
the reference / the functional oracle (tests/test_graph_vs_oracle.py)."""
import torch
import torch.nn.functional as F


def run_graph(graph, sd, x, train=False, bf16=False, stop_after=None, relu_masks=None):
    """relu_masks: optional {tensor name: bool mask}; when given, ReLU of that unit is applied as y*mask so that a
    backward comparison is not dominated by sign flips of near-zero pre-activations (bf16 forward noise)."""
    def r(t):
        return t.to(torch.bfloat16).to(torch.float32) if bf16 else t

    T = {"in": x}
    for u in graph.units:
        src = [T[s] for s in u.srcs]
        h = src[0] if len(src) == 1 else torch.cat(src, 1)
        if u.up:
            h = F.interpolate(h, scale_factor=2, mode="bilinear", align_corners=(u.up == 2))
        w = sd[u.conv + ".weight"]
        if u.transposed:
            y = F.conv_transpose2d(r(h), r(w), None, u.stride, u.pad)
        else:
            if u.reflect:
                h = F.pad(h, (u.pad,) * 4, mode="reflect")
                y = F.conv2d(r(h), r(w), None, u.stride, 0)
            else:
                y = F.conv2d(r(h), r(w), None, u.stride, u.pad)
        if u.bn:
            g, b = sd[u.bn + ".weight"], sd[u.bn + ".bias"]
            if train:
                mean = y.mean((0, 2, 3))
                var = y.var((0, 2, 3), unbiased=False)
                ys = y.to(torch.float16).to(torch.float32) if bf16 else y   # device stores the raw conv output as fp16
                y = (ys - mean[None, :, None, None]) * torch.rsqrt(var + 1e-5)[None, :, None, None]
                y = y * g[None, :, None, None] + b[None, :, None, None]
            else:
                y = F.batch_norm(y, sd[u.bn + ".running_mean"], sd[u.bn + ".running_var"], g, b, False, 0.1, 1e-5)
        if u.relu:
            if relu_masks is not None and u.out in relu_masks:
                y = y * relu_masks[u.out].to(y.dtype)
            else:
                y = F.relu(y)
        if u.resid:
            y = y + T[u.resid]
        if u.tanh:
            y = torch.tanh(y)
        T[u.out] = y
        if stop_after is not None and u.out == stop_after:
            break
    return T
