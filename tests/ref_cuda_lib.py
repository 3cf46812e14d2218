"""ctypes access to oracle/_ref/libref_cuda{,_fma}.so: the REFERENCE's own CUDA kernels + the host half of its GPU
MultiProposalTarget operator, cut out of the reference tree and compiled by oracle/build_ref_cuda.py (test
infrastructure only).  Device-pointer arguments take torch CUDA tensors; host-pointer arguments numpy arrays."""
import ctypes
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_LIBS = {}
P, I, F = ctypes.c_void_p, ctypes.c_int, ctypes.c_float


def path(fma=False):
    return os.path.join(ROOT, "oracle", "_ref", "libref_cuda_fma.so" if fma else "libref_cuda.so")


def available(fma=False):
    return os.path.exists(path(fma))


def lib(fma=False):
    if fma not in _LIBS:
        _LIBS[fma] = ctypes.CDLL(path(fma))
    return _LIBS[fma]


def _d(t):
    return P(0 if t is None else t.data_ptr())


def _h(a):
    return a.ctypes.data_as(P)


def generate_anchors(feat_stride, scales, ratios, fma=False):
    s, r = np.asarray(scales, np.float32), np.asarray(ratios, np.float32)
    out = np.zeros((len(r) * len(s), 4), np.float32)
    n = lib(fma).ref_generate_anchors(I(feat_stride), _h(r), I(len(r)), _h(s), I(len(s)), _h(out))
    assert n == out.size
    return out


def host_assign(gt_boxes, rois, valid_ranges, post=300, fma=False):
    """multi_proposal_target.cu:435-578 on host arrays; gt_boxes [B,100,5]; rois [B*post,5] (modified copy returned)."""
    gt = np.ascontiguousarray(gt_boxes, np.float32)
    assert gt.shape[1] == 100, "the reference hard-codes 100 GT rows per image"
    B = gt.shape[0]
    rois = np.ascontiguousarray(rois, np.float32).copy()
    vr = np.ascontiguousarray(valid_ranges, np.float32)
    labels = np.zeros(B * post, np.float32)
    bt = np.zeros((B * post, 4), np.float32)
    bw = np.zeros((B * post, 4), np.float32)
    lib(fma).ref_mpt_host_assign(_h(gt), _h(rois), _h(labels), _h(bt), _h(bw), _h(vr), I(B), I(post))
    return dict(rois=rois, label=labels, bbox_target=bt, bbox_weight=bw)


def get_props(deltas, im_info, anchors, scores, valid_ranges, B, A, H, W, stride=16, fma=False):
    import torch
    boxes = torch.zeros(B * A * H * W, 6, device="cuda")
    rc = lib(fma).ref_get_props(_d(boxes), _d(deltas), _d(im_info), _d(anchors), _d(scores), _d(valid_ranges), I(B), I(A),
                                I(H), I(W), I(stride))
    assert rc == 0
    return boxes


def nms(dets, B, A, H, W, post=300, fma=False):
    """dets [B*A*H*W, 6] CUDA tensor (permuted in place, as the reference does); returns propsout [B*post, 5]."""
    import torch
    out = torch.zeros(B * post, 5, device="cuda")
    rc = lib(fma).ref_nms(_d(dets), I(post), I(B), I(A), I(W), I(H), _d(out))
    assert rc == 0
    return out


def expf(x, fma=False):
    import torch
    y = torch.empty_like(x)
    assert lib(fma).ref_expf(_d(x), _d(y), I(x.numel())) == 0
    return y


def dpsroi_fwd(data, rois, trans, spatial_scale, output_dim, group_size, pooled, part_size, spp, trans_std, fma=False):
    """data [B,C,H,W], rois [N,5], trans [N,2*ncls,part,part] or None -> (out [N,output_dim,P,P], top_count)."""
    import torch
    Bn, C, H, W = data.shape
    N = rois.shape[0]
    no_trans = trans is None
    ncls = 1 if no_trans else trans.shape[1] // 2
    cec = output_dim if no_trans else output_dim // ncls
    out = torch.zeros(N, output_dim, pooled, pooled, device="cuda")
    cnt = torch.zeros_like(out)
    rc = lib(fma).ref_dpsroi_fwd(I(out.numel()), _d(data), F(spatial_scale), I(C), I(H), I(W), I(pooled), I(pooled),
                                 _d(rois), _d(trans), I(int(no_trans)), F(trans_std), I(spp), I(output_dim),
                                 I(group_size), I(part_size), I(ncls), I(cec), _d(out), _d(cnt))
    assert rc == 0
    return out, cnt


def dpsroi_bwd(top_diff, top_count, data, rois, trans, spatial_scale, output_dim, group_size, pooled, part_size, spp,
               trans_std, fma=False):
    import torch
    Bn, C, H, W = data.shape
    N = rois.shape[0]
    no_trans = trans is None
    ncls = 1 if no_trans else trans.shape[1] // 2
    cec = output_dim if no_trans else output_dim // ncls
    dd = torch.zeros_like(data)
    td = None if no_trans else torch.zeros_like(trans)
    rc = lib(fma).ref_dpsroi_bwd(I(top_diff.numel()), _d(top_diff), _d(top_count), I(N), F(spatial_scale), I(C), I(H),
                                 I(W), I(pooled), I(pooled), I(output_dim), _d(dd), _d(td), _d(data), _d(rois),
                                 _d(trans), I(int(no_trans)), F(trans_std), I(spp), I(group_size), I(part_size),
                                 I(ncls), I(cec))
    assert rc == 0
    return dd, td


def psroi_fwd(data, rois, spatial_scale, output_dim, group_size, pooled, fma=False):
    import torch
    Bn, C, H, W = data.shape
    out = torch.zeros(rois.shape[0], output_dim, pooled, pooled, device="cuda")
    rc = lib(fma).ref_psroi_fwd(I(out.numel()), _d(data), F(spatial_scale), I(C), I(H), I(W), I(pooled), I(pooled),
                                _d(rois), I(output_dim), I(group_size), _d(out))
    assert rc == 0
    return out


def psroi_bwd(top_diff, rois, data_shape, spatial_scale, output_dim, group_size, pooled, fma=False):
    import torch
    Bn, C, H, W = data_shape
    dd = torch.zeros(data_shape, device="cuda")
    rc = lib(fma).ref_psroi_bwd(I(top_diff.numel()), _d(top_diff), I(rois.shape[0]), F(spatial_scale), I(C), I(H), I(W),
                                I(pooled), I(pooled), I(group_size), I(output_dim), _d(dd), _d(rois))
    assert rc == 0
    return dd


def deform_im2col(data, offset, kh=3, kw=3, pad=2, stride=1, dil=2, dgroups=4, fma=False):
    """data [N,C,H,W], offset [N,dg*2*kh*kw,Hc,Wc] -> col [N, C*kh*kw, Hc, Wc] (per-image launches like the operator)."""
    import torch
    N, C, H, W = data.shape
    Hc, Wc = offset.shape[2], offset.shape[3]
    col = torch.zeros(N, C * kh * kw, Hc, Wc, device="cuda")
    for n in range(N):
        rc = lib(fma).ref_deform_im2col(_d(data[n]), _d(offset[n]), I(C), I(H), I(W), I(kh), I(kw), I(pad), I(stride),
                                        I(dil), I(dgroups), I(Hc), I(Wc), _d(col[n]))
        assert rc == 0
    return col


def deform_col2im(col, data, offset, kh=3, kw=3, pad=2, stride=1, dil=2, dgroups=4, fma=False):
    """Gradients of deform_im2col: (grad_im [N,C,H,W], grad_offset like offset)."""
    import torch
    N, C, H, W = data.shape
    Hc, Wc = offset.shape[2], offset.shape[3]
    gi = torch.zeros_like(data)
    go = torch.zeros_like(offset)
    for n in range(N):
        a = (I(C), I(H), I(W), I(kh), I(kw), I(pad), I(stride), I(dil), I(dgroups), I(Hc), I(Wc))
        assert lib(fma).ref_deform_col2im(_d(col[n]), _d(offset[n]), *a, _d(gi[n])) == 0
        assert lib(fma).ref_deform_col2im_coord(_d(col[n]), _d(data[n]), _d(offset[n]), *a, _d(go[n])) == 0
    return gi, go
