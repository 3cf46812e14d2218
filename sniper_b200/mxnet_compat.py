"""`import mxnet as mx` for the reference's SYMBOL FILES: the slice of MXNet's symbolic front end that
symbols/faster/resnet_mx_101_e2e.py and symbols/faster/mobilenetv2_e2e.py touch, so that those files run UNCHANGED and
the graph they describe can be inspected (names, shapes, `-symbol.json`) and bound to this package's executor.

    from sniper_b200 import mxnet_compat
    mod = mxnet_compat.load_symbol_file('/path/to/SNIPER/symbols/faster/resnet_mx_101_e2e.py')
    inst = mod.resnet_mx_101_e2e(n_proposals=400, momentum=0.995)
    sym = inst.get_symbol_rcnn(config)              # the reference's own code builds the graph, node by node
    inst.infer_shape(data_shape_dict)               # symbols/symbol.py:41-47 -> Symbol.infer_shape below
    sym.save('prefix-symbol.json')                  # MXNet 1.2 wire format
    net = sym.bind('cuda:0')                        # -> model.SniperResNet101 (only for the graphs this package executes)

What is restated here (host code only, no kernels), each from the MXNet sources vendored in the reference tree:
  * naming: NameManager (python/mxnet/name.py:24-66), generated creators (python/mxnet/symbol/register.py:111-176:
    attribute values are `str(v)`, `dtype` goes through np.dtype().name, hint = lower-cased function name),
    hidden keys lr_mult / wd_mult / ... -> `__lr_mult__` (src/c_api/c_api_symbolic.cc:40-45,121-141);
  * composition: missing inputs become variables `<node>_<arg>` that inherit the node's attribute dictionary
    (3rdparty/nnvm/src/core/symbolic.cc:51-58,283-341), BatchNorm tags its moving_mean / moving_var with `__init__`
    (src/operator/nn/batch_norm.cc:614-623);
  * listing: arguments / auxiliary states / outputs in post-order DFS (symbolic.cc ListInputNames / ListOutputNames);
  * `tojson()` / `load_json()`: 3rdparty/nnvm/src/pass/saveload_json.cc:67-88 (node), :133-147 (graph), :206-243
    (post-order numbering, arg_nodes, node_row_ptr over ALL outputs incl. hidden ones, heads [node, index, version]);
    pinned by the MXNet-1.2 file the reference tree ships (tests/python/mkl/data/test_mkldnn_test_mkldnn_model_model1.json);
  * shape inference of the 21 operators the two symbol files use (table OPS below; each entry cites its InferShape).
Nothing here computes: `bind()` hands the recognised graph to `model.SniperResNet101`, anything else raises.
"""
import collections
import importlib.util
import json
import os
import sys
import types

import numpy as np

MXNET_VERSION = 10200          # include/mxnet/base.h MXNET_MAJOR 1, MINOR 2, PATCH 0 of the vendored fork
HIDDEN_KEYS = ("ctx_group", "lr_mult", "wd_mult", "force_mirroring", "mirror_stage")


# ------------------------------------------------------------------------------------------------ naming / attribute scope
class NameManager(object):
    current = None

    def __init__(self):
        self._counter = {}
        self._old = None

    def get(self, name, hint):
        if name:
            return name
        n = self._counter.get(hint, 0)
        self._counter[hint] = n + 1
        return "%s%d" % (hint, n)

    def __enter__(self):
        self._old = NameManager.current
        NameManager.current = self
        return self

    def __exit__(self, *a):
        NameManager.current = self._old


NameManager.current = NameManager()


class AttrScope(object):
    current = None

    def __init__(self, **kwargs):
        self._attr = {k: str(v) for k, v in kwargs.items()}
        self._old = None

    def get(self, attr):
        if self._attr:
            ret = dict(self._attr)
            if attr:
                ret.update(attr)
            return ret
        return dict(attr) if attr else {}

    def __enter__(self):
        self._old = AttrScope.current
        merged = dict(self._old._attr)
        merged.update(self._attr)
        self._attr = merged
        AttrScope.current = self
        return self

    def __exit__(self, *a):
        AttrScope.current = self._old


AttrScope.current = AttrScope()


# ------------------------------------------------------------------------------------------------ attribute parsing
def _tuple(v, n=None, default=None):
    if v is None:
        return default
    if isinstance(v, str):
        v = v.strip().strip("()[]")
        v = tuple(int(float(x)) for x in v.split(",") if x.strip())
    elif isinstance(v, (int, np.integer)):
        v = (int(v),)
    v = tuple(int(x) for x in v)
    if n is not None and len(v) == 1 and n > 1:
        v = v * n
    return v


def _bool(v, default=False):
    if v is None:
        return default
    if isinstance(v, str):
        return v.strip().lower() in ("true", "1")
    return bool(v)


def _int(v, default=None):
    return default if v is None else int(float(v))


# ------------------------------------------------------------------------------------------------ operator table
# name -> dict(inputs=fn(attrs) -> names, aux=indices of mutable inputs, nout=total outputs, nvis=visible outputs,
#              outs=fn(attrs) -> output names, shape=fn(attrs, in_shapes) -> (in_shapes, out_shapes))
def _conv_out(h, k, s, p, d):
    return (h + 2 * p - (d * (k - 1) + 1)) // s + 1


def _shape_conv(a, ins, deform=False):
    """nn/convolution.cc ConvolutionShape (2-D NCHW); contrib/deformable_convolution-inl.h:356-442."""
    data = ins[0]
    k = _tuple(a["kernel"])
    s = _tuple(a.get("stride"), 2, (1, 1)) or (1, 1)
    d = _tuple(a.get("dilate"), 2, (1, 1)) or (1, 1)
    p = _tuple(a.get("pad"), 2, (0, 0)) or (0, 0)
    nf, ng = _int(a["num_filter"]), _int(a.get("num_group"), 1)
    if data is None:
        return ins, [None]
    N, C, H, W = data
    wi = 2 if deform else 1
    ins = list(ins)
    ins[wi] = (nf, C // ng, k[0], k[1])
    if not _bool(a.get("no_bias")):
        ins[wi + 1] = (nf,)
    Ho, Wo = _conv_out(H, k[0], s[0], p[0], d[0]), _conv_out(W, k[1], s[1], p[1], d[1])
    if deform:
        ins[1] = (N, 2 * k[0] * k[1] * _int(a.get("num_deformable_group"), 1), Ho, Wo)
    return ins, [(N, nf, Ho, Wo)]


def _shape_bn(a, ins):
    """nn/batch_norm.cc BatchNormShape: per-channel vectors on `axis` (default 1)."""
    data = ins[0]
    if data is None:
        return ins, [None, None, None]
    c = (data[_int(a.get("axis"), 1)],)
    return [data, c, c, c, c], [data, c, c]


def _shape_same(a, ins):
    return ins, [ins[0]]


def _shape_binary(a, ins):
    s = ins[0] if ins[0] is not None else ins[1]
    return [s, s], [s]


def _shape_pool(a, ins):
    """nn/pooling.cc PoolingShape: `valid` floors, `full` ceils, global_pool -> 1x1."""
    data = ins[0]
    if data is None:
        return ins, [None]
    N, C, H, W = data
    if _bool(a.get("global_pool")):
        return ins, [(N, C, 1, 1)]
    k = _tuple(a["kernel"])
    s = _tuple(a.get("stride"), 2, (1, 1)) or (1, 1)
    p = _tuple(a.get("pad"), 2, (0, 0)) or (0, 0)
    if a.get("pooling_convention", "valid") == "full":
        f = lambda h, kk, ss, pp: 1 + -(-(h + 2 * pp - kk) // ss)
    else:
        f = lambda h, kk, ss, pp: 1 + (h + 2 * pp - kk) // ss
    return ins, [(N, C, f(H, k[0], s[0], p[0]), f(W, k[1], s[1], p[1]))]


def _shape_concat(a, ins):
    dim = _int(a.get("dim"), 1)
    if any(s is None for s in ins):
        return ins, [None]
    out = list(ins[0])
    out[dim] = sum(s[dim] for s in ins)
    return ins, [tuple(out)]


def _shape_reshape(a, ins):
    """tensor/matrix_op-inl.h InferReshapeShape: 0 copy, -1 infer, -2 copy the rest, -3 merge two, -4 split."""
    src = ins[0]
    if src is None:
        return ins, [None]
    spec = list(_tuple(a["shape"]))
    if _bool(a.get("reverse")):
        raise NotImplementedError("Reshape(reverse=True)")
    out, i, j, infer = [], 0, 0, -1
    while j < len(spec):
        v = spec[j]
        if v == 0:
            out.append(src[i]); i += 1
        elif v == -1:
            infer = len(out); out.append(1); i += 1
        elif v == -2:
            out += list(src[i:]); i = len(src)
        elif v == -3:
            out.append(src[i] * src[i + 1]); i += 2
        elif v == -4:
            d1, d2 = spec[j + 1], spec[j + 2]
            j += 2
            d0 = src[i]; i += 1
            if d1 == -1:
                d1 = d0 // d2
            if d2 == -1:
                d2 = d0 // d1
            out += [d1, d2]
        else:
            out.append(v); i += 1
        j += 1
    total = int(np.prod(src))
    if infer >= 0:
        rest = int(np.prod(out))
        out[infer] = total // rest if rest else 0
    assert int(np.prod(out)) == total, "Reshape: %s -> %s" % (src, out)
    return ins, [tuple(out)]


def _shape_fc(a, ins):
    """nn/fully_connected.cc FullyConnectedShape (flatten=True)."""
    data = ins[0]
    if data is None:
        return ins, [None]
    nh = _int(a["num_hidden"])
    ins = list(ins)
    ins[1] = (nh, int(np.prod(data[1:])))
    if not _bool(a.get("no_bias")):
        ins[2] = (nh,)
    return ins, [(data[0], nh)]


def _shape_softmax_output(a, ins):
    """softmax_output-inl.h:281-325: label = data without the channel axis (multi_output), or (N,); a label given by
    the caller in the flattened (N, rest) form is accepted as MXNet accepts it."""
    data, label = ins
    if data is None:
        return ins, [None]
    if label is None:
        if _bool(a.get("multi_output")):
            label = (data[0],) + tuple(data[2:])
        else:
            label = (data[0],)
    return [data, label], [data]


def _shape_mpt(a, ins):
    """multi_proposal_target-inl.h:109-132."""
    c = ins[0]
    if c is None:
        return ins, [None] * 4
    n = c[0] * _int(a.get("rpn_post_nms_top_n"), 300)
    return ins, [(n, 5), (n, 1), (n, 4), (n, 4)]


def _shape_mp(a, ins):
    """multi_proposal-inl.h:109-123."""
    c = ins[0]
    if c is None:
        return ins, [None] * 2
    n = c[0] * _int(a.get("rpn_post_nms_top_n"), 300)
    return ins, [(n, 5), (n,)]


def _shape_dpsroi(a, ins):
    """contrib/deformable_psroi_pooling-inl.h:214-246."""
    rois = ins[1]
    if rois is None:
        return ins, [None, None]
    s = (rois[0], _int(a["output_dim"]), _int(a["pooled_size"]), _int(a["pooled_size"]))
    return ins, [s, s]


def _conv_inputs(a):
    return ["data", "weight"] + ([] if _bool(a.get("no_bias")) else ["bias"])


def _one(names):
    return lambda a: list(names)


OPS = {
    "Convolution": dict(inputs=_conv_inputs, shape=_shape_conv),
    "_contrib_DeformableConvolution": dict(
        inputs=lambda a: ["data", "offset", "weight"] + ([] if _bool(a.get("no_bias")) else ["bias"]),
        shape=lambda a, i: _shape_conv(a, i, deform=True)),
    "BatchNorm": dict(inputs=_one(["data", "gamma", "beta", "moving_mean", "moving_var"]), aux=(3, 4), nout=3,
                      nvis=lambda a: 3 if _bool(a.get("output_mean_var")) else 1, outs=_one(["output", "mean", "var"]),
                      shape=_shape_bn),
    "Activation": dict(inputs=_one(["data"]), shape=_shape_same),
    "Cast": dict(inputs=_one(["data"]), shape=_shape_same),
    "clip": dict(inputs=_one(["data"]), shape=_shape_same),
    "BlockGrad": dict(inputs=_one(["data"]), shape=_shape_same),
    "MakeLoss": dict(inputs=_one(["data"]), shape=_shape_same),
    "smooth_l1": dict(inputs=_one(["data"]), shape=_shape_same),
    "SoftmaxActivation": dict(inputs=_one(["data"]), shape=_shape_same),
    "Flatten": dict(inputs=_one(["data"]),
                    shape=lambda a, i: (i, [None if i[0] is None else (i[0][0], int(np.prod(i[0][1:])))])),
    "Dropout": dict(inputs=_one(["data"]), nout=2, nvis=lambda a: 1, outs=_one(["output", "mask"]),
                    shape=lambda a, i: (i, [i[0], i[0]])),
    "Pooling": dict(inputs=_one(["data"]), shape=_shape_pool, nvis=lambda a: 1,
                    nout=lambda a: 2 if (MKLDNN_BUILD and a.get("pool_type", "max") == "max") else 1),
    "Concat": dict(inputs=None, shape=_shape_concat, key_var_num_args="num_args"),
    "Reshape": dict(inputs=_one(["data"]), shape=_shape_reshape),
    "FullyConnected": dict(inputs=_conv_inputs, shape=_shape_fc),
    "SoftmaxOutput": dict(inputs=_one(["data", "label"]), shape=_shape_softmax_output),
    "elemwise_add": dict(inputs=_one(["lhs", "rhs"]), shape=_shape_binary),
    "elemwise_sub": dict(inputs=_one(["lhs", "rhs"]), shape=_shape_binary),
    "elemwise_mul": dict(inputs=_one(["lhs", "rhs"]), shape=_shape_binary),
    # crowd_boxes: mobilenetv2_e2e.py:260 passes it, the vendored operator (multi_proposal_target-inl.h:164-166) does not
    # list it -- accepted as an optional trailing input so that the file composes (this package's kernel ignores crowds,
    # as the vendored operator does)
    "MultiProposalTarget": dict(
        inputs=lambda a: ["cls_prob", "bbox_pred", "im_info", "gt_boxes", "valid_ranges"] +
        (["crowd_boxes"] if a.get("__has_crowd__") else []),
        nout=4, outs=_one(["rois", "label", "bbox_target", "bbox_weight"]), shape=_shape_mpt),
    "MultiProposal": dict(inputs=_one(["cls_prob", "bbox_pred", "im_info"]), nout=2, outs=_one(["output", "score"]),
                          shape=_shape_mp),
    "_contrib_DeformablePSROIPooling": dict(
        inputs=lambda a: ["data", "rois"] + ([] if _bool(a.get("no_trans")) else ["trans"]),
        nout=2, nvis=lambda a: 1, outs=_one(["output", "top_count"]), shape=_shape_dpsroi),
}


MKLDNN_BUILD = False     # nn/pooling.cc GetNumOutputs: an MKLDNN build gives max Pooling a hidden workspace output.  The
                         # SNIPER build (CUDA, no MKLDNN) has one; the MXNet-1.2 sample file the wire format is pinned
                         # against was written by an MKLDNN build, so its test flips this switch.


def _op_nout(op, attrs=None):
    n = OPS[op].get("nout", 1)
    return n(attrs or {}) if callable(n) else n


# ------------------------------------------------------------------------------------------------ graph
class _Node(object):
    __slots__ = ("op", "name", "attrs", "inputs")

    def __init__(self, op, name, attrs=None, inputs=None):
        self.op, self.name = op, name
        self.attrs = dict(attrs or {})          # str -> str, as nnvm keeps them
        self.inputs = list(inputs or [])        # [(node, index)]

    @property
    def is_variable(self):
        return self.op is None

    def num_outputs(self):
        return 1 if self.op is None else _op_nout(self.op, self.attrs)

    def parsed(self):
        return {k: v for k, v in self.attrs.items()}


def _dfs(heads):
    """Post-order DFS from the head entries, inputs in order (nnvm/include/nnvm/graph.h PostOrderDFSVisit)."""
    seen, order = set(), []
    for h, _ in heads:
        if id(h) in seen:
            continue
        stack = [(h, 0)]
        seen.add(id(h))
        while stack:
            node, i = stack.pop()
            if i < len(node.inputs):
                stack.append((node, i + 1))
                child = node.inputs[i][0]
                if id(child) not in seen:
                    seen.add(id(child))
                    stack.append((child, 0))
            else:
                order.append(node)
    return order


class Symbol(object):
    """python/mxnet/symbol/symbol.py Symbol: a list of output entries of a shared graph."""

    def __init__(self, heads):
        self._heads = list(heads)

    # ---- composition helpers
    def __iter__(self):
        return (self[i] for i in range(len(self._heads)))

    def __len__(self):
        return len(self._heads)

    def __getitem__(self, i):
        if isinstance(i, str):
            names = self.list_outputs()
            i = names.index(i)
        return Symbol([self._heads[i]])

    def _entry(self):
        if len(self._heads) != 1:
            raise ValueError("Argument is a tuple, single value is required")
        return self._heads[0]

    @property
    def name(self):
        return self._heads[0][0].name if len(self._heads) == 1 else None

    def attr(self, key):
        n = self._entry()[0]
        if key in n.attrs:
            return n.attrs[key]
        if key in HIDDEN_KEYS:
            return n.attrs.get("__%s__" % key)
        return None

    def list_attr(self):
        return dict(self._entry()[0].attrs)

    def attr_dict(self):
        return {n.name: dict(n.attrs) for n in _dfs(self._heads) if n.attrs}

    def _set_attr(self, **kwargs):
        n = self._entry()[0]
        for k, v in kwargs.items():
            n.attrs["__%s__" % k if k in HIDDEN_KEYS else k] = str(v)

    def get_internals(self):
        heads = []
        for n in _dfs(self._heads):
            nvis = 1 if n.is_variable else _visible(n)
            heads += [(n, i) for i in range(nvis)]
        return Symbol(heads)

    def get_children(self):
        n = self._entry()[0]
        return Symbol(list(n.inputs)) if n.inputs else None

    # ---- arithmetic (symbol.py __add__/__sub__/__mul__ -> _internal._Plus/_Minus/_Mul)
    def __add__(self, o):
        return _create("elemwise_add", "_plus", [self, o], {}, None)

    def __sub__(self, o):
        return _create("elemwise_sub", "_minus", [self, o], {}, None)

    def __mul__(self, o):
        return _create("elemwise_mul", "_mul", [self, o], {}, None)

    __radd__ = __add__
    __rmul__ = __mul__

    # ---- listing
    def _aux_ids(self, order):
        aux = set()
        for n in order:
            if not n.is_variable:
                for i in OPS[n.op].get("aux", ()):
                    if i < len(n.inputs) and n.inputs[i][0].is_variable:
                        aux.add(id(n.inputs[i][0]))
        return aux

    def list_arguments(self):
        order = _dfs(self._heads)
        aux = self._aux_ids(order)
        return [n.name for n in order if n.is_variable and id(n) not in aux]

    def list_auxiliary_states(self):
        order = _dfs(self._heads)
        aux = self._aux_ids(order)
        return [n.name for n in order if n.is_variable and id(n) in aux]

    def list_inputs(self):
        return [n.name for n in _dfs(self._heads) if n.is_variable]

    def list_outputs(self):
        out = []
        for n, i in self._heads:
            if n.is_variable:
                out.append(n.name)
                continue
            outs = OPS[n.op].get("outs")
            r = outs(n.attrs)[i] if outs else ("output" if _op_nout(n.op, n.attrs) == 1 else "output%d" % i)
            out.append(n.name + "_" + r if n.name else r)
        return out

    # ---- shapes (src/executor/infer_graph_attr_pass.cc: here one forward sweep in topological order is enough, every
    # operator of the two symbol files derives its parameter shapes from its data input)
    def _infer(self, known):
        order = _dfs(self._heads)
        shapes = {}
        for n in order:
            if n.is_variable:
                s = known.get(n.name)
                if s is None and "__shape__" in n.attrs:
                    s = _tuple(n.attrs["__shape__"])
                shapes[(id(n), 0)] = None if s is None else tuple(int(x) for x in s)
                continue
            ins = [shapes.get((id(m), i)) for m, i in n.inputs]
            new_ins, outs = OPS[n.op]["shape"](n.attrs, ins)
            for (m, i), s_old, s_new in zip(n.inputs, ins, new_ins):
                if s_new is None:
                    continue
                if s_old is not None and tuple(s_old) != tuple(s_new) and int(np.prod(s_old)) != int(np.prod(s_new)):
                    raise ValueError("infer_shape: %s input %s: provided %s, inferred %s" % (n.name, m.name, s_old, s_new))
                if s_old is None:
                    shapes[(id(m), i)] = tuple(s_new)
            for i, s in enumerate(outs):
                shapes[(id(n), i)] = None if s is None else tuple(s)
        return order, shapes

    def infer_shape(self, *args, **kwargs):
        if args:
            kwargs = dict(zip(self.list_arguments(), args))
        order, shapes = self._infer(kwargs)
        aux = self._aux_ids(order)
        arg_s = [shapes[(id(n), 0)] for n in order if n.is_variable and id(n) not in aux]
        aux_s = [shapes[(id(n), 0)] for n in order if n.is_variable and id(n) in aux]
        out_s = [shapes[(id(n), i)] for n, i in self._heads]
        if any(s is None for s in arg_s + aux_s + out_s):
            return None, None, None           # MXNet: incomplete inference
        return arg_s, out_s, aux_s

    def infer_shape_partial(self, **kwargs):
        order, shapes = self._infer(kwargs)
        aux = self._aux_ids(order)
        f = lambda s: () if s is None else s
        return ([f(shapes[(id(n), 0)]) for n in order if n.is_variable and id(n) not in aux],
                [f(shapes[(id(n), i)]) for n, i in self._heads],
                [f(shapes[(id(n), 0)]) for n in order if n.is_variable and id(n) in aux])

    # ---- wire format
    def _json_obj(self):
        order = _dfs(self._heads)
        nid = {id(n): k for k, n in enumerate(order)}
        nodes, arg_nodes, row = [], [], [0]
        for k, n in enumerate(order):
            d = collections.OrderedDict()
            d["op"] = "null" if n.is_variable else n.op
            d["name"] = n.name
            public = {a: v for a, v in n.attrs.items() if a != "__has_crowd__"}
            if public:
                d["attrs"] = collections.OrderedDict(sorted(public.items()))
            d["inputs"] = [[nid[id(m)], i, 0] for m, i in n.inputs]
            nodes.append(d)
            if n.is_variable:
                arg_nodes.append(k)
            row.append(row[-1] + n.num_outputs())
        g = collections.OrderedDict()
        g["nodes"] = nodes
        g["arg_nodes"] = arg_nodes
        g["node_row_ptr"] = row
        g["heads"] = [[nid[id(n)], i, 0] for n, i in self._heads]
        g["attrs"] = {"mxnet_version": ["int", MXNET_VERSION]}
        return g

    def tojson(self):
        return json.dumps(self._json_obj(), indent=2)

    def save(self, fname):
        with open(fname, "w") as f:
            f.write(self.tojson())

    def debug_str(self):
        return "\n".join("%s %s(%s)" % (n.name, n.op or "Variable", ", ".join(m.name for m, _ in n.inputs))
                         for n in _dfs(self._heads))

    # ---- executor
    def bind(self, device="cuda:0", **overrides):
        """The graphs this package executes are recognised by their parameter set; anything else is refused."""
        from . import symbols
        return symbols.bind_graph(self, device=device, **overrides)


def _visible(n):
    nv = OPS[n.op].get("nvis")
    return nv(n.attrs) if nv else _op_nout(n.op, n.attrs)


def Variable(name, attr=None, shape=None, lr_mult=None, wd_mult=None, dtype=None, init=None, stype=None, **kwargs):
    """symbol.py var(): attributes are stored as `__key__` strings."""
    if not isinstance(name, str):
        raise TypeError("Expect a string for variable `name`")
    attrs = AttrScope.current.get(attr)
    if shape is not None:
        attrs["__shape__"] = str(tuple(shape))
    if lr_mult is not None:
        attrs["__lr_mult__"] = str(lr_mult)
    if wd_mult is not None:
        attrs["__wd_mult__"] = str(wd_mult)
    if dtype is not None:
        attrs["__dtype__"] = str({"float32": 0, "float64": 1, "float16": 2, "uint8": 3, "int32": 4, "int8": 5,
                                  "int64": 6}[np.dtype(dtype).name])
    if init is not None:
        attrs["__init__"] = init if isinstance(init, str) else init.dumps()
    for k, v in kwargs.items():
        if k.startswith("__") and k.endswith("__"):
            attrs[k] = str(v)
        else:
            raise ValueError("Attribute name=%s is not supported. Additional attributes must start and end with "
                             "double underscores, e.g, __yourattr__" % k)
    return Symbol([(_Node(None, name, attrs), 0)])


var = Variable


def Group(symbols):
    heads = []
    for s in symbols:
        if not isinstance(s, Symbol):
            raise TypeError("Expected a list of symbols as input")
        heads += s._heads
    return Symbol(heads)


def _create(op, hint, sym_args, kwargs, name, attr=None):
    """register.py generated creator + MXSymbolCreateAtomicSymbol + nnvm Symbol::Compose."""
    spec = OPS[op]
    kwargs = dict(kwargs)
    kwargs.update(AttrScope.current.get(attr))
    name = NameManager.current.get(name, hint)
    sym_kwargs, attrs = {}, {}
    for k, v in kwargs.items():
        if isinstance(v, Symbol):
            sym_kwargs[k] = v
        elif v is not None:
            if k == "dtype":
                v = np.dtype(v).name
            attrs[k] = str(v)
    kv = spec.get("key_var_num_args")
    if kv and kv not in attrs:
        attrs[kv] = str(len(sym_args) + len(sym_kwargs))
    for k in list(attrs):
        for h in HIDDEN_KEYS:
            if k == h or (k.startswith(h) and k.rfind(h) == 0):
                attrs["__%s__" % k] = attrs.pop(k)
                break
            if k.endswith(h):
                raise ValueError("setting variable attributes with %s is deprecated" % k)
    if op == "MultiProposalTarget" and "crowd_boxes" in sym_kwargs:
        attrs["__has_crowd__"] = "1"
    node = _Node(op, name, attrs)
    if spec["inputs"] is None:                       # variable-length (Concat): positional only
        if sym_kwargs:
            raise ValueError("Variable length function do not accept kwargs")
        node.inputs = [s._entry() for s in sym_args]
        return Symbol([(node, i) for i in range(_visible(node))])
    arg_names = spec["inputs"](attrs)
    if len(sym_args) > len(arg_names):
        raise ValueError("Incorrect number of arguments, requires %d, provided %d" % (len(arg_names), len(sym_args)))
    inputs, matched = [], 0
    for i, an in enumerate(arg_names):
        if i < len(sym_args):
            inputs.append(sym_args[i]._entry())
        elif an in sym_kwargs:
            inputs.append(sym_kwargs[an]._entry())
            matched += 1
        else:
            v = _Node(None, (name + "_" + an) if name else an, attrs)      # inherits the parent's attributes
            v.attrs.pop("__has_crowd__", None)
            inputs.append((v, 0))
    if matched != len(sym_kwargs):
        bad = [k for k in sym_kwargs if k not in arg_names[len(sym_args):]]
        raise ValueError("Symbol.Compose: Keyword argument name %s not found. Candidate arguments: %s"
                         % (", ".join(bad), ", ".join(arg_names[len(sym_args):])))
    node.inputs = inputs
    if op == "BatchNorm":                            # FSetInputVarAttrOnCompose, batch_norm.cc:614-623
        for i, val in ((3, '["zero", {}]'), (4, '["one", {}]')):
            v = inputs[i][0]
            if v.is_variable and "__init__" not in v.attrs:
                v.attrs["__init__"] = val
    return Symbol([(node, i) for i in range(_visible(node))])


def _make_creator(op, fname):
    def creator(*args, **kwargs):
        name = kwargs.pop("name", None)
        attr = kwargs.pop("attr", None)
        kwargs.pop("out", None)
        for a in args:
            if not isinstance(a, Symbol):
                raise TypeError("Positional arguments must be Symbol instances, but got %s" % str(a))
        return _create(op, fname.lower(), list(args), kwargs, name, attr)
    creator.__name__ = fname
    return creator


def load_json(json_str):
    """nnvm LoadJSON (saveload_json.cc:158-203); `attr` / `param` of pre-1.0 files are merged as the loader does."""
    g = json.loads(json_str)
    nodes = []
    for jn in g["nodes"]:
        attrs = {}
        for key in ("attrs", "attr", "param"):
            attrs.update(jn.get(key, {}))
        op = None if jn["op"] == "null" else jn["op"]
        if op is not None and op not in OPS:
            raise ValueError("Failed loading Op %s of type %s: not one of the operators of the SNIPER symbols"
                             % (jn["name"], op))
        nodes.append(_Node(op, jn["name"], attrs, [(nodes[e[0]], e[1]) for e in jn["inputs"]]))
    for k in g["arg_nodes"]:
        assert nodes[k].is_variable
    return Symbol([(nodes[e[0]], e[1]) for e in g["heads"]])


def load(fname):
    with open(fname) as f:
        return load_json(f.read())


# ------------------------------------------------------------------------------------------------ mx.nd / mx.random / mx.model
class NDArray(np.ndarray):
    """The handful of NDArray methods the symbol files use on parameters (init_weight_rcnn, checkpoint_callback)."""

    def asnumpy(self):
        return np.asarray(self)

    @property
    def context(self):
        return "cpu(0)"


def _nd(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(NDArray)


_rng = np.random.RandomState(0)


def _nd_zeros(shape, ctx=None, dtype=None, **kw):
    return _nd(np.zeros(shape, np.float32))


def _nd_ones(shape, ctx=None, dtype=None, **kw):
    return _nd(np.ones(shape, np.float32))


def _nd_array(a, ctx=None, dtype=None):
    return _nd(np.asarray(a))


def _random_normal(loc=0, scale=1, shape=None, ctx=None, dtype=None, **kw):
    return _nd(_rng.normal(loc, scale, size=shape))


def _random_uniform(low=0, high=1, shape=None, ctx=None, dtype=None, **kw):
    return _nd(_rng.uniform(low, high, size=shape))


def _random_seed(s):
    global _rng
    _rng = np.random.RandomState(int(s))


def save_checkpoint(prefix, epoch, symbol, arg_params, aux_params):
    """python/mxnet/model.py:366-394: `prefix-symbol.json` + `prefix-%04d.params` with arg:/aux: key prefixes."""
    from . import checkpoint as ck
    if symbol is not None:
        symbol.save("%s-symbol.json" % prefix)
    ck.write_params("%s-%04d.params" % (prefix, epoch),
                    {k: np.asarray(v) for k, v in arg_params.items()}, {k: np.asarray(v) for k, v in aux_params.items()})


def load_checkpoint(prefix, epoch):
    """python/mxnet/model.py:397-429."""
    from . import checkpoint as ck
    arg, aux = ck.read_params("%s-%04d.params" % (prefix, epoch))
    return load("%s-symbol.json" % prefix), {k: _nd(v) for k, v in arg.items()}, {k: _nd(v) for k, v in aux.items()}


# ------------------------------------------------------------------------------------------------ module objects
def _namespace(_modname, **members):
    m = types.ModuleType(_modname)
    for k, v in members.items():
        setattr(m, k, v)
    return m


_PUBLIC = {   # python-visible creator name -> operator
    "Convolution": "Convolution", "BatchNorm": "BatchNorm", "Activation": "Activation", "Cast": "Cast", "clip": "clip",
    "BlockGrad": "BlockGrad", "stop_gradient": "BlockGrad", "MakeLoss": "MakeLoss", "make_loss": "MakeLoss",
    "smooth_l1": "smooth_l1", "SoftmaxActivation": "SoftmaxActivation", "Pooling": "Pooling", "Concat": "Concat",
    "concat": "Concat", "Reshape": "Reshape", "reshape": "Reshape", "FullyConnected": "FullyConnected",
    "SoftmaxOutput": "SoftmaxOutput", "Softmax": "SoftmaxOutput", "elemwise_add": "elemwise_add",
    "elemwise_sub": "elemwise_sub", "elemwise_mul": "elemwise_mul", "Flatten": "Flatten", "flatten": "Flatten",
    "Dropout": "Dropout", "MultiProposalTarget": "MultiProposalTarget", "MultiProposal": "MultiProposal",
}
_CONTRIB = {"DeformableConvolution": "_contrib_DeformableConvolution",
            "DeformablePSROIPooling": "_contrib_DeformablePSROIPooling",
            "MultiProposalTarget": "MultiProposalTarget", "MultiProposal": "MultiProposal"}


def _clip(data=None, a_min=None, a_max=None, name=None, attr=None, **kw):
    """`mx.sym.clip(data, 0, 6, name=...)` -- a_min / a_max are positional in the generated signature."""
    return _create("clip", "clip", [], dict(data=data, a_min=a_min, a_max=a_max, **kw), name, attr)


def _build_module():
    creators = {k: _make_creator(op, k) for k, op in _PUBLIC.items()}
    creators["clip"] = _clip
    contrib_creators = {k: _make_creator(op, k) for k, op in _CONTRIB.items()}
    contrib_sym = _namespace("mxnet.contrib.symbol", **contrib_creators)
    common = dict(Variable=Variable, var=Variable, Group=Group, Symbol=Symbol, load=load, load_json=load_json,
                  contrib=contrib_sym, **creators)
    sym = _namespace("mxnet.symbol", **common)
    contrib = _namespace("mxnet.contrib", sym=contrib_sym, symbol=contrib_sym)
    nd = _namespace("mxnet.ndarray", zeros=_nd_zeros, ones=_nd_ones, array=_nd_array, NDArray=NDArray)
    random = _namespace("mxnet.random", normal=_random_normal, uniform=_random_uniform, seed=_random_seed)
    nd.random = random
    model = _namespace("mxnet.model", save_checkpoint=save_checkpoint, load_checkpoint=load_checkpoint)
    name = _namespace("mxnet.name", NameManager=NameManager)
    attribute = _namespace("mxnet.attribute", AttrScope=AttrScope)
    from . import operator_py
    mx = _namespace("mxnet", sym=sym, symbol=sym, contrib=contrib, nd=nd, ndarray=nd, random=random, model=model,
                    name=name, attribute=attribute, operator=operator_py, AttrScope=AttrScope, NameManager=NameManager,
                    cpu=lambda i=0: "cpu(%d)" % i, gpu=lambda i=0: "cuda:%d" % i, __version__="1.2.0")
    return mx


mx = _build_module()
sym = symbol = mx.sym
contrib = mx.contrib
nd = mx.nd
random = mx.random
model = mx.model


def install():
    """Registers this module's `mx` object as `mxnet` (only if no real MXNet is importable / imported)."""
    for k, m in (("mxnet", mx), ("mxnet.symbol", mx.sym), ("mxnet.sym", mx.sym), ("mxnet.contrib", mx.contrib),
                 ("mxnet.ndarray", mx.nd), ("mxnet.nd", mx.nd), ("mxnet.random", mx.random), ("mxnet.model", mx.model)):
        sys.modules.setdefault(k, m)
    return mx


def load_symbol_file(path, module_name=None):
    """Executes one of the reference's symbol files (symbols/faster/*.py) unchanged against this front end.  The three
    imports at its head resolve to: `mxnet` -> this module; `symbols.symbol` -> the file next to it in the reference tree
    (pure numpy) or, if absent, sniper_b200.symbols; `operator_py.box_annotator_ohem` -> sniper_b200.operator_py (the OHEM
    operator itself is out of scope; the e2e symbols only import it)."""
    install()
    path = os.path.abspath(path)
    root = os.path.dirname(os.path.dirname(os.path.dirname(path)))       # <repo>/symbols/faster/x.py -> <repo>
    saved = {k: sys.modules.get(k) for k in ("symbols", "symbols.symbol", "operator_py", "operator_py.box_annotator_ohem")}
    try:
        base = os.path.join(root, "symbols", "symbol.py")
        pkg = types.ModuleType("symbols")
        pkg.__path__ = [os.path.join(root, "symbols")]
        sys.modules["symbols"] = pkg
        if os.path.exists(base):
            spec = importlib.util.spec_from_file_location("symbols.symbol", base)
            m = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(m)
        else:
            from . import symbols as m
        sys.modules["symbols.symbol"] = m
        from . import operator_py
        opkg = types.ModuleType("operator_py")
        opkg.__path__ = []
        sys.modules["operator_py"] = opkg
        sys.modules["operator_py.box_annotator_ohem"] = operator_py
        spec = importlib.util.spec_from_file_location(module_name or os.path.splitext(os.path.basename(path))[0], path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        return mod
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
