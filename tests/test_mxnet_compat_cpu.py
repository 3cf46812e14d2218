"""`sniper_b200.mxnet_compat`: the MXNet symbolic front end the reference's symbol files need.

* wire format / naming / composition pinned by the MXNet-1.2 `-symbol.json` the reference tree ships
  (SNIPER-mxnet/tests/python/mkl/data/test_mkldnn_test_mkldnn_model_model1.json): the graph is REPLAYED through the
  creators from its operator nodes alone and must serialise to the same JSON (auto-created variables, inherited
  attributes, post-order numbering, arg_nodes, node_row_ptr, heads);
* the reference's own symbol files (resnet_mx_101_e2e.py, mobilenetv2_e2e.py) execute unchanged and give the committed
  fixture tests/golden/ref_symbols.json (regenerated here when /root/reference exists);
* the fixture pins `sniper_b200.symbols.NetSymbol` (names, shapes, order of arguments / auxiliary states / outputs) and
  `recognise_graph`.
"""
import hashlib
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
GOLD = os.path.join(ROOT, "tests", "golden", "ref_symbols.json")
SAMPLE = os.path.join(REF, "SNIPER-mxnet/tests/python/mkl/data/test_mkldnn_test_mkldnn_model_model1.json")
needs_ref = pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not present")


def _replay(j):
    """Rebuilds a graph from the OPERATOR nodes of a symbol JSON: variables that a node would create by itself
    (`<node>_<arg>`) are left to Compose, every other variable is created explicitly."""
    from sniper_b200 import mxnet_compat as MC
    syms = {}
    with MC.NameManager():
        for k, n in enumerate(j["nodes"]):
            if n["op"] == "null":
                continue
            ins = []
            names = MC.OPS[n["op"]]["inputs"](n.get("attrs", {}))
            kwargs = dict(n.get("attrs", {}))
            for an, (src, idx, _) in zip(names, n["inputs"]):
                sn = j["nodes"][src]
                if sn["op"] == "null":
                    if sn["name"] == n["name"] + "_" + an:
                        continue                                  # auto-created by Compose
                    if src not in syms:
                        syms[src] = MC.Variable(sn["name"])
                    kwargs[an] = syms[src]
                else:
                    kwargs[an] = syms[src][idx]
            syms[k] = getattr(MC.sym, n["op"])(name=n["name"], **kwargs)
    heads = [syms[h[0]][h[1]] for h in j["heads"]]
    return heads[0] if len(heads) == 1 else MC.Group(heads)


@needs_ref
def test_wire_format_against_the_mxnet12_file_in_the_reference_tree(monkeypatch):
    from sniper_b200 import mxnet_compat as MC
    monkeypatch.setattr(MC, "MKLDNN_BUILD", True)      # the sample was saved by an MKLDNN build (max Pooling: 2 outputs)
    text = open(SAMPLE).read()
    j = json.loads(text)
    assert j["attrs"] == {"mxnet_version": ["int", 10200]}
    # (1) load -> save reproduces the file, object for object
    s = MC.load_json(text)
    assert json.loads(s.tojson()) == j
    # (2) composing the same network through the creators reproduces it too
    r = _replay(j)
    assert json.loads(r.tojson()) == j
    assert r.list_arguments() == [j["nodes"][k]["name"] for k in j["arg_nodes"]]
    assert r.list_outputs() == ["softmax_output"] or r.list_outputs()[0].endswith("_output")
    # (3) shapes of a VGG-16 on 224x224
    arg, out, aux = r.infer_shape(data=(2, 3, 224, 224))
    d = dict(zip(r.list_arguments(), arg))
    assert d["conv1_1_weight"] == (64, 3, 3, 3) and d["conv5_3_bias"] == (512,) and not aux


def test_naming_and_composition_rules():
    from sniper_b200 import mxnet_compat as MC
    mx = MC.mx
    with MC.NameManager():
        data = mx.sym.Variable("data")
        c = mx.sym.Convolution(data=data, num_filter=8, kernel=(3, 3), pad=(1, 1), no_bias=True, name="c1")
        b = mx.sym.BatchNorm(data=c, fix_gamma=False, eps=2e-5, name="bn")
        p = mx.symbol.Pooling(data=b, kernel=(3, 3), stride=(2, 2), pad=(1, 1), pool_type="max")
        q = mx.sym.Cast(data=p, dtype=np.float16)
        f = mx.sym.FullyConnected(data=q, num_hidden=5, lr_mult=0.01, name="fc")
        s = f + f
        s2 = s * f - f
    assert p.name == "pooling0" and q.name == "cast0" and s.name == "_plus0" and s2.name == "_minus0"
    assert s2.list_arguments() == ["data", "c1_weight", "bn_gamma", "bn_beta", "fc_weight", "fc_bias"]
    assert s2.list_auxiliary_states() == ["bn_moving_mean", "bn_moving_var"]
    assert s2.list_outputs() == ["_minus0_output"]
    j = json.loads(s2.tojson())
    byname = {n["name"]: n for n in j["nodes"]}
    assert byname["cast0"]["attrs"] == {"dtype": "float16"}
    assert byname["c1"]["attrs"] == {"kernel": "(3, 3)", "no_bias": "True", "num_filter": "8", "pad": "(1, 1)"}
    assert byname["c1_weight"]["attrs"] == byname["c1"]["attrs"]               # inherited (symbolic.cc:312-313)
    assert byname["fc"]["attrs"] == {"__lr_mult__": "0.01", "num_hidden": "5"}   # hidden key (c_api_symbolic.cc:121-141)
    assert byname["bn_moving_mean"]["attrs"]["__init__"] == '["zero", {}]'
    assert byname["bn_moving_var"]["attrs"]["__init__"] == '["one", {}]'
    assert f.attr("lr_mult") == "0.01"
    # BatchNorm has three outputs of which one is visible: node_row_ptr counts all of them
    k = [n["name"] for n in j["nodes"]].index("bn")
    assert j["node_row_ptr"][k + 1] - j["node_row_ptr"][k] == 3
    arg, out, aux = s2.infer_shape(data=(2, 3, 16, 16))
    assert dict(zip(s2.list_arguments(), arg))["fc_weight"] == (5, 8 * 8 * 8) and out == [(2, 5)] and aux == [(8,), (8,)]
    # keyword mismatch is an error, as in nnvm
    with pytest.raises(ValueError):
        mx.sym.Activation(dat=data, act_type="relu")
    # round trip
    assert json.loads(MC.load_json(s2.tojson()).tojson()) == j
    # Reshape codes
    r = mx.sym.Reshape(data=mx.sym.Variable("x"), shape=(0, 2, -1, 0))
    assert r.infer_shape(x=(4, 42, 32, 32))[1] == [(4, 2, 672, 32)]
    r = mx.sym.Reshape(data=mx.sym.Variable("x"), shape=(-3, -2))
    assert r.infer_shape(x=(4, 5, 6, 7))[1] == [(20, 6, 7)]
    r = mx.sym.Reshape(data=mx.sym.Variable("x"), shape=(-4, 2, -1, -2))
    assert r.infer_shape(x=(4, 5, 6))[1] == [(2, 2, 5, 6)]


def test_save_checkpoint_writes_symbol_json_and_params(tmp_path):
    from sniper_b200 import checkpoint as ck
    from sniper_b200 import mxnet_compat as MC
    mx = MC.mx
    with MC.NameManager():
        net = mx.sym.FullyConnected(data=mx.sym.Variable("data"), num_hidden=3, name="bbox_pred")
    arg = {"bbox_pred_weight": mx.random.normal(0, 0.01, shape=(3, 4)), "bbox_pred_bias": mx.nd.zeros(shape=(3,))}
    prefix = str(tmp_path / "e2e")
    mx.model.save_checkpoint(prefix, 7, net, arg, {})
    sym, a, x = mx.model.load_checkpoint(prefix, 7)
    assert sym.list_arguments() == ["data", "bbox_pred_weight", "bbox_pred_bias"] and not x
    assert np.array_equal(a["bbox_pred_weight"].asnumpy(), np.asarray(arg["bbox_pred_weight"]))
    assert set(ck.read_params(prefix + "-0007.params")[0]) == set(arg)
    assert (arg["bbox_pred_weight"].T * mx.nd.array([1, 2, 3])).T.shape == (3, 4)          # checkpoint_callback's arithmetic


@needs_ref
def test_reference_symbol_files_execute_unchanged_and_match_the_fixture():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import run_ref_symbols as R
    got = json.loads(json.dumps(R.build(), sort_keys=True))
    want = json.load(open(GOLD))
    assert got == want, "tests/golden/ref_symbols.json is stale: python oracle/run_ref_symbols.py"


def _gold():
    return json.load(open(GOLD))


@pytest.mark.parametrize("tag", ["resnet101_train_fp32", "resnet101_train_fp16"])
def test_netsymbol_equals_the_graph_the_reference_builds(tag):
    from types import SimpleNamespace
    from sniper_b200 import symbols
    g = _gold()[tag]
    cfg = SimpleNamespace(dataset=SimpleNamespace(NUM_CLASSES=81), network=SimpleNamespace(NUM_ANCHORS=21),
                          TRAIN=SimpleNamespace(AUTO_FOCUS=False, fp16=g["cfg"]["fp16"]))
    inst = symbols.resnet_mx_101_e2e(n_proposals=400, momentum=0.995)
    sym = inst.get_symbol_rcnn(cfg)
    data = {n: tuple(s) for n, s in g["arguments"] if n in sym.data_names()}
    arg, out, aux = sym.infer_shape(**data)
    ref_args = {n: tuple(s) for n, s in g["arguments"]}
    assert dict(zip(sym.list_arguments(), map(tuple, arg))) == ref_args
    assert dict(zip(sym.list_auxiliary_states(), map(tuple, aux))) == {n: tuple(s) for n, s in g["auxiliary"]}
    assert list(zip(sym.list_outputs(), map(tuple, out))) == [(n, tuple(s)) for n, s in g["outputs"]]
    # parameters in the reference's DFS order (what `arg_params` iteration / `.params` files follow)
    params_ref = [n for n, _ in g["arguments"] if n not in data]
    params_ours = [n for n in sym.list_arguments() if n not in data]
    assert sorted(params_ref) == sorted(params_ours)
    assert [n for n, _ in g["auxiliary"]] == sym.list_auxiliary_states()


@needs_ref
def test_recognise_graph_accepts_the_resnet_graphs_and_refuses_others():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import run_ref_symbols as R
    from sniper_b200 import mxnet_compat as MC
    from sniper_b200 import symbols
    res = MC.load_symbol_file(os.path.join(REF, "symbols/faster/resnet_mx_101_e2e.py"))
    cfg = R.load_config("sniper_res101_e2e.yml")
    cfg.TRAIN.BATCH_IMAGES = 20
    with MC.NameManager():
        sym = res.resnet_mx_101_e2e(n_proposals=400, momentum=0.995).get_symbol_rcnn(cfg)
    info = symbols.recognise_graph(sym)
    assert info == dict(batch_images=20, num_anchors=21, num_classes=81, bf16=bool(cfg.TRAIN.fp16), is_train=True,
                        autofocus=False)
    # through the wire format too
    assert symbols.recognise_graph(MC.load_json(sym.tojson())) == info
    assert hashlib.md5(sym.tojson().encode()).hexdigest() == _gold()[
        "resnet101_train_fp16" if cfg.TRAIN.fp16 else "resnet101_train_fp32"]["json_md5"] or cfg.TRAIN.BATCH_IMAGES != 20
    # the reference's own init_weight_rcnn / check_parameter_shapes run on the graph's shapes
    inst = res.resnet_mx_101_e2e(n_proposals=400, momentum=0.995)
    with MC.NameManager():
        inst.get_symbol_rcnn(cfg)
    shapes = R.train_shapes(cfg, 20, 16)
    inst.infer_shape(shapes)
    arg, aux = {}, {}
    inst.init_weight_rcnn(cfg, arg, aux)
    assert arg["rpn_conv_3x3_weight"].shape == (512, 3072, 3, 3) and not np.asarray(arg["offset_weight"]).any()
    assert abs(float(np.asarray(arg["fc_new_1_weight"]).std()) - 0.01) < 1e-3
    # the reference's own checkpoint_callback (:6-17) writes -symbol.json + .params with the *_test copies
    import tempfile
    from sniper_b200 import checkpoint as ck
    with tempfile.TemporaryDirectory() as d:
        prefix = os.path.join(d, "e2e")
        small = {"bbox_pred_weight": np.asarray(arg["bbox_pred_weight"]), "bbox_pred_bias": np.asarray(arg["bbox_pred_bias"]) + 1}
        res.checkpoint_callback(inst.get_bbox_param_names(), prefix, None, None)(3, inst.symbol, small, {})
        a2, x2 = ck.read_params(prefix + "-0004.params")
        assert set(a2) == {"bbox_pred_weight", "bbox_pred_bias", "bbox_pred_weight_test", "bbox_pred_bias_test"} and not x2
        assert np.allclose(a2["bbox_pred_bias_test"], small["bbox_pred_bias"] * np.array([0.1, 0.1, 0.2, 0.2]))
        assert set(small) == {"bbox_pred_weight", "bbox_pred_bias"}
        back = MC.load(prefix + "-symbol.json")
        assert back.list_arguments() == inst.symbol.list_arguments() and back.tojson() == inst.symbol.tojson()
    # the RPN-only graphs: the test graph is recognised (executor: forward_rpn), the training graph is refused
    with MC.NameManager():
        rpn_test = res.resnet_mx_101_e2e(test_nbatch=2).get_symbol_rpn(cfg, is_train=False)
    assert symbols.recognise_graph(rpn_test)["rpn_only"] is True and symbols.recognise_graph(rpn_test)["batch_images"] == 2
    with MC.NameManager():
        rpn_train = res.resnet_mx_101_e2e().get_symbol_rpn(cfg)
    with pytest.raises(NotImplementedError):
        symbols.recognise_graph(rpn_train)
    # a ResNet-50 graph is not ours
    r50 = MC.load_symbol_file(os.path.join(REF, "symbols/faster/resnet_mx_50_e2e.py"))
    with MC.NameManager():
        s50 = r50.resnet_mx_50_e2e(n_proposals=400, momentum=0.995).get_symbol_rcnn(cfg)
    with pytest.raises(NotImplementedError):
        symbols.recognise_graph(s50)


def test_mobilenet_symbol_equals_the_graph_the_reference_builds():
    from types import SimpleNamespace
    from sniper_b200 import symbols
    g = _gold()["mobilenetv2_train"]
    cfg = SimpleNamespace(dataset=SimpleNamespace(NUM_CLASSES=81), network=SimpleNamespace(NUM_ANCHORS=15))
    inst = symbols.mobilenetv2_e2e()
    sym = inst.get_symbol_rcnn(cfg)
    data = {n: tuple(s) for n, s in g["arguments"] if n in sym.data_names()}
    assert set(data) == set(sym.data_names())
    arg, out, aux = sym.infer_shape(**data)
    assert dict(zip(sym.list_arguments(), map(tuple, arg))) == {n: tuple(s) for n, s in g["arguments"]}
    assert list(zip(sym.list_auxiliary_states(), map(tuple, aux))) == [(n, tuple(s)) for n, s in g["auxiliary"]]
    assert list(zip(sym.list_outputs(), map(tuple, out))) == [(n, tuple(s)) for n, s in g["outputs"]]
    inst.infer_shape(data)
    a, x = {}, {}
    inst.init_weight_rcnn(cfg, a, x, seed=0)
    assert a["fc_new_1_weight"].shape == (512, 12544) and not a["offset_weight"].any()


@needs_ref
def test_recognise_graph_accepts_the_mobilenet_training_graph():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import run_ref_symbols as R
    from sniper_b200 import mxnet_compat as MC
    from sniper_b200 import symbols
    mob = MC.load_symbol_file(os.path.join(REF, "symbols/faster/mobilenetv2_e2e.py"))
    cfg = R.load_config("sniper_mobilenetv2_e2e.yml")
    cfg.TRAIN.BATCH_IMAGES = 40
    with MC.NameManager():
        sym = mob.mobilenetv2_e2e(n_proposals=400, momentum=0.995).get_symbol_rcnn(cfg)
    info = symbols.recognise_graph(sym)
    assert info == dict(batch_images=40, num_anchors=15, num_classes=81, bf16=True, is_train=True, autofocus=False,
                        network="mobilenetv2")
    with MC.NameManager():
        test_sym = mob.mobilenetv2_e2e(test_nbatch=2).get_symbol_rcnn(cfg, is_train=False)
    with pytest.raises(NotImplementedError):
        symbols.recognise_graph(test_sym)


@needs_ref
def test_reference_custom_operator_file_registers_on_the_shim():
    """lib/operator_py/box_annotator_ohem.py (the reference's Python custom operator, written against mx.operator) loads
    unchanged: its `@mx.operator.register` lands in sniper_b200.operator_py's registry and its CustomOpProp answers the
    shape / argument queries (the OHEM operator's arithmetic itself is outside the hot path)."""
    import importlib.util
    from sniper_b200 import mxnet_compat as MC
    from sniper_b200 import operator_py
    MC.install()
    spec = importlib.util.spec_from_file_location("ref_box_annotator_ohem",
                                                  os.path.join(REF, "lib/operator_py/box_annotator_ohem.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    assert "BoxAnnotatorOHEM" in operator_py._REGISTRY and issubclass(m.BoxAnnotatorOHEMProp, operator_py.CustomOpProp)
    prop = operator_py._REGISTRY["BoxAnnotatorOHEM"](num_classes="81", num_reg_classes="1", roi_per_img="300")
    assert prop.list_arguments() == ['cls_score', 'bbox_pred', 'labels', 'bbox_targets', 'bbox_weights']
    ins, outs = prop.infer_shape([(16, 300, 81), (16, 300, 4), (16, 300), (16, 300, 4), (16, 300, 4)])[:2]
    assert tuple(outs[0]) == (16, 300) and len(outs) == len(prop.list_outputs())
    assert isinstance(prop.create_operator(None, ins, None), operator_py.CustomOp)


@pytest.mark.parametrize("is_train", [True, False])
def test_rpn_symbol_equals_the_graph_the_reference_builds(is_train):
    from types import SimpleNamespace
    from sniper_b200 import symbols
    g = _gold()["resnet101_rpn_train" if is_train else "resnet101_rpn_test"]
    cfg = SimpleNamespace(network=SimpleNamespace(NUM_ANCHORS=21))
    inst = symbols.resnet_mx_101_e2e()
    sym = inst.get_symbol_rpn(cfg, is_train=is_train)
    data = {n: tuple(s) for n, s in g["arguments"] if n in sym.data_names()}
    assert set(data) == set(sym.data_names())
    arg, out, aux = sym.infer_shape(**data)
    assert dict(zip(sym.list_arguments(), map(tuple, arg))) == {n: tuple(s) for n, s in g["arguments"]}
    assert list(zip(sym.list_auxiliary_states(), map(tuple, aux))) == [(n, tuple(s)) for n, s in g["auxiliary"]]
    assert list(zip(sym.list_outputs(), map(tuple, out))) == [(n, tuple(s)) for n, s in g["outputs"]]
    inst.infer_shape(data)
    a = {}
    inst.init_weight_rpn(cfg, a, {}, seed=0)
    assert set(a) == {n + s for n in ("rpn_conv_3x3", "rpn_cls_score", "rpn_bbox_pred", "stage4_unit1_offset",
                                      "stage4_unit2_offset", "stage4_unit3_offset") for s in ("_weight", "_bias")}
