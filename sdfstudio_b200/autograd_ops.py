"""Differentiable forms of the weights / compositing operators (training path).

The reference trains through ATen autograd over rays.py:131-230 and renderers.py:42-295; here every op keeps its forward
kernel and gets an explicit backward kernel (csrc/render_backward.cu: sdfb200_render_backward, sdfb200_weights_backward).
Used automatically by rays.py / renderers.py when an input requires grad; the no-grad (rendering) path does not go through here.
"""
import torch

from . import _lib


def needs_grad(*ts) -> bool:
    return torch.is_grad_enabled() and any(t is not None and torch.is_tensor(t) and t.requires_grad for t in ts)


class WeightsFromAlphasFn(torch.autograd.Function):
    """alphas [R,S] -> weights [R,S], transmittance [R,S+1] (rays.py:194-230).  Gradients flow back through the weights and
    through every transmittance column (bg_transmittance = transmittance[:, -1], models/neus.py:101)."""

    @staticmethod
    def forward(ctx, alphas):
        lib = _lib.load()
        a = _lib.f32c(alphas)
        R, S = a.shape
        w = torch.empty_like(a)
        T = torch.empty(R, S + 1, device=a.device, dtype=torch.float32)
        _lib.check(lib.sdfb200_weights_from_alphas(_lib.ptr(a), R, S, _lib.ptr(w), _lib.ptr(T), _lib.stream_ptr()), "sdfb200_weights_from_alphas")
        ctx.save_for_backward(a)
        return w, T

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g_w, g_T):
        lib = _lib.load()
        (a,) = ctx.saved_tensors
        R, S = a.shape
        g_w = _lib.f32c(g_w) if g_w is not None else torch.zeros_like(a)
        g_T = _lib.f32c(g_T) if g_T is not None else None
        g_a = torch.empty_like(a)
        _lib.check(lib.sdfb200_weights_backward(_lib.ptr(a), None, 0, R, S, _lib.ptr(g_w), _lib.ptr(g_T), S + 1, _lib.ptr(g_a), _lib.stream_ptr()),
                   "sdfb200_weights_backward")
        return g_a


class WeightsFromDensityFn(torch.autograd.Function):
    """densities [R,S], euclidean bins [R,S+1] -> weights [R,S], transmittance [R,S] (rays.py:131-192).  Gradients flow to the
    densities through the weights AND the transmittance (VolSDF composites the background model with transmittance[:, -1],
    models/volsdf.py:67-68 + base_surface_model.py:329)."""

    @staticmethod
    def forward(ctx, density, bins):
        lib = _lib.load()
        d = _lib.f32c(density)
        R, S = d.shape
        w = torch.empty_like(d)
        T = torch.empty_like(d)
        _lib.check(lib.sdfb200_weights_from_density(_lib.ptr(d), _lib.ptr(bins), R, S, _lib.ptr(w), _lib.ptr(T), _lib.stream_ptr()),
                   "sdfb200_weights_from_density")
        ctx.save_for_backward(d, bins)
        return w, T

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g_w, g_T):
        lib = _lib.load()
        d, bins = ctx.saved_tensors
        R, S = d.shape
        g_w = _lib.f32c(g_w) if g_w is not None else torch.zeros_like(d)
        g_T = _lib.f32c(g_T) if g_T is not None else None
        g_d = torch.empty_like(d)
        _lib.check(lib.sdfb200_weights_backward(_lib.ptr(d), _lib.ptr(bins), 1, R, S, _lib.ptr(g_w), _lib.ptr(g_T), S, _lib.ptr(g_d), _lib.stream_ptr()),
                   "sdfb200_weights_backward")
        return g_d, None


def _render_backward(ctx_t, bg_mode, grads, g_weights_in, want_rgb_s, want_nrm_s):
    """shared by RenderFn / RenderAlphasFn: per-ray output gradients -> (g_weights, g_rgb_samples, g_normal_samples)."""
    lib = _lib.load()
    w, rgb, nrm, bins, bg, acc, depth = ctx_t
    R, S = w.shape
    g_rgb, g_depth, g_nrm, g_acc = grads
    g_rgb = _lib.f32c(g_rgb) if (g_rgb is not None and rgb is not None) else None
    g_depth = _lib.f32c(g_depth) if (g_depth is not None and bins is not None) else None
    g_nrm = _lib.f32c(g_nrm) if (g_nrm is not None and nrm is not None) else None
    g_acc = _lib.f32c(g_acc) if g_acc is not None else None
    g_w = torch.empty_like(w)
    g_rgb_s = torch.empty(R, S, 3, device=w.device, dtype=torch.float32) if (want_rgb_s and rgb is not None) else None
    g_nrm_s = torch.empty(R, S, 3, device=w.device, dtype=torch.float32) if (want_nrm_s and nrm is not None) else None
    _lib.check(lib.sdfb200_render_backward(_lib.ptr(w), _lib.ptr(rgb), _lib.ptr(nrm), _lib.ptr(bins), _lib.ptr(bg), bg_mode, R, S, _lib.ptr(acc),
                                           _lib.ptr(depth), _lib.ptr(g_rgb), _lib.ptr(g_depth), _lib.ptr(g_nrm), _lib.ptr(g_acc),
                                           _lib.ptr(g_weights_in), _lib.ptr(g_w), _lib.ptr(g_rgb_s), _lib.ptr(g_nrm_s), _lib.stream_ptr()),
               "sdfb200_render_backward")
    return g_w, g_rgb_s, g_nrm_s


class RenderFn(torch.autograd.Function):
    """weights [R,S] (+ rgb, normals [R,S,3], bins [R,S+1]) -> rgb [R,3], UNCLIPPED expected depth [R], normal [R,3],
    accumulation [R], steps_minmax [2] (sdfb200_render).  Absent inputs are passed as None and yield zero-filled outputs."""

    @staticmethod
    def forward(ctx, weights, rgb, normals, bins, bg, bg_mode):
        lib = _lib.load()
        w = _lib.f32c(weights)
        R, S = w.shape
        dev = w.device
        rgb = _lib.f32c(rgb) if rgb is not None else None
        normals = _lib.f32c(normals) if normals is not None else None
        o_rgb = torch.zeros(R, 3, device=dev)
        o_depth = torch.zeros(R, device=dev)
        o_nrm = torch.zeros(R, 3, device=dev)
        o_acc = torch.empty(R, device=dev)
        mm = torch.tensor([float("inf"), float("-inf")], device=dev)
        out = _lib.RenderOut()
        out.accumulation = o_acc.data_ptr()
        if rgb is not None:
            out.rgb = o_rgb.data_ptr()
        if normals is not None:
            out.normal = o_nrm.data_ptr()
        if bins is not None:
            out.depth, out.steps_minmax = o_depth.data_ptr(), mm.data_ptr()
        _lib.check(lib.sdfb200_render(_lib.ptr(w), _lib.ptr(rgb), _lib.ptr(normals), _lib.ptr(bins), _lib.ptr(bg), bg_mode, 0, 0, R, S, out,
                                      _lib.stream_ptr()), "sdfb200_render")
        ctx.tensors = (w, rgb, normals, bins, bg, o_acc, o_depth)
        ctx.bg_mode = bg_mode
        ctx.mark_non_differentiable(mm)
        return o_rgb, o_depth, o_nrm, o_acc, mm

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g_rgb, g_depth, g_nrm, g_acc, _g_mm):
        g_w, g_rgb_s, g_nrm_s = _render_backward(ctx.tensors, ctx.bg_mode, (g_rgb, g_depth, g_nrm, g_acc), None,
                                                 ctx.needs_input_grad[1], ctx.needs_input_grad[2])
        return g_w, g_rgb_s, g_nrm_s, None, None, None


class RenderAlphasFn(torch.autograd.Function):
    """alphas [R,S] -> weights, rgb, UNCLIPPED depth, normal, accumulation, bg_transmittance, steps_minmax in one forward launch
    (sdfb200_render_alphas); backward = sdfb200_render_backward -> sdfb200_weights_backward."""

    @staticmethod
    def forward(ctx, alphas, rgb, normals, bins, bg, bg_mode):
        lib = _lib.load()
        a = _lib.f32c(alphas)
        R, S = a.shape
        dev = a.device
        rgb, normals = _lib.f32c(rgb), _lib.f32c(normals)
        w = torch.empty(R, S, device=dev)
        o_rgb, o_depth, o_nrm = torch.empty(R, 3, device=dev), torch.empty(R, device=dev), torch.empty(R, 3, device=dev)
        o_acc, o_bgT = torch.empty(R, device=dev), torch.empty(R, device=dev)
        mm = torch.tensor([float("inf"), float("-inf")], device=dev)
        out = _lib.RenderOut()
        out.rgb, out.depth, out.normal, out.accumulation, out.steps_minmax = (o_rgb.data_ptr(), o_depth.data_ptr(), o_nrm.data_ptr(), o_acc.data_ptr(),
                                                                               mm.data_ptr())
        _lib.check(lib.sdfb200_render_alphas(_lib.ptr(a), _lib.ptr(rgb), _lib.ptr(normals), _lib.ptr(bins), _lib.ptr(bg), bg_mode, 0, R, S, _lib.ptr(w),
                                             o_bgT.data_ptr(), out, _lib.stream_ptr()), "sdfb200_render_alphas")
        ctx.tensors = (w, rgb, normals, bins, bg, o_acc, o_depth)
        ctx.alphas = a
        ctx.bg_mode = bg_mode
        ctx.mark_non_differentiable(mm)
        return w, o_rgb, o_depth, o_nrm, o_acc, o_bgT, mm

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g_w_in, g_rgb, g_depth, g_nrm, g_acc, g_bgT, _g_mm):
        lib = _lib.load()
        a = ctx.alphas
        R, S = a.shape
        g_w_in = _lib.f32c(g_w_in) if g_w_in is not None else None
        g_w, g_rgb_s, g_nrm_s = _render_backward(ctx.tensors, ctx.bg_mode, (g_rgb, g_depth, g_nrm, g_acc), g_w_in,
                                                 ctx.needs_input_grad[1], ctx.needs_input_grad[2])
        g_a = None
        if ctx.needs_input_grad[0]:
            g_a = torch.empty_like(a)
            g_last = _lib.f32c(g_bgT) if g_bgT is not None else None
            _lib.check(lib.sdfb200_weights_backward(_lib.ptr(a), None, 0, R, S, _lib.ptr(g_w), _lib.ptr(g_last), 1, _lib.ptr(g_a), _lib.stream_ptr()),
                       "sdfb200_weights_backward")
        return g_a, g_rgb_s, g_nrm_s, None, None, None
