"""TEST INFRASTRUCTURE (oracle/): executes the reference's own symbol files -- symbols/faster/resnet_mx_101_e2e.py and
symbols/faster/mobilenetv2_e2e.py, unchanged, with the reference's own yml configs -- against sniper_b200.mxnet_compat
and writes what the resulting graphs look like (argument / auxiliary / output names and shapes, operator census, the
MD5 of the `-symbol.json` text) to tests/golden/ref_symbols.json.  The committed fixture lets the GPU box (which has no
/root/reference) check `sniper_b200.symbols` against the graphs the reference's code builds.

Rewrites done in memory only (nothing of the reference is copied): `easydict` (absent here) is a 10-line attribute dict;
`yaml.load(f)` gets the Loader argument PyYAML >= 6 insists on.

    python oracle/run_ref_symbols.py            # needs /root/reference
"""
import hashlib
import importlib.util
import json
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("SNIPER_REFERENCE", "/root/reference")
sys.path.insert(0, ROOT)


class edict(dict):
    def __init__(self, d=None, **kw):
        super().__init__()
        for k, v in dict(d or {}, **kw).items():
            self[k] = v

    def __setitem__(self, k, v):
        if isinstance(v, dict) and not isinstance(v, edict):
            v = edict(v)
        super().__setitem__(k, v)

    __setattr__ = __setitem__

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)


def load_config(yml):
    """configs/faster/default_configs.py `config` after update_config(yml), fresh for every call."""
    import yaml
    sys.modules["easydict"] = types.SimpleNamespace(EasyDict=edict)
    orig = yaml.load
    yaml.load = lambda f, Loader=None: orig(f, Loader=Loader or yaml.FullLoader)
    try:
        spec = importlib.util.spec_from_file_location("ref_default_configs", os.path.join(REF, "configs/faster/default_configs.py"))
        m = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(m)
        m.update_config(os.path.join(REF, "configs/faster", yml))
        return m.config
    finally:
        yaml.load = orig
        sys.modules.pop("easydict", None)


def describe(sym, data_shapes):
    arg_s, out_s, aux_s = sym.infer_shape(**data_shapes)
    assert arg_s is not None, "incomplete shape inference"
    ops = {}
    for n in json.loads(sym.tojson())["nodes"]:
        ops[n["op"]] = ops.get(n["op"], 0) + 1
    return dict(arguments=[[n, list(s)] for n, s in zip(sym.list_arguments(), arg_s)],
                auxiliary=[[n, list(s)] for n, s in zip(sym.list_auxiliary_states(), aux_s)],
                outputs=[[n, list(s)] for n, s in zip(sym.list_outputs(), out_s)],
                ops=ops, json_md5=hashlib.md5(sym.tojson().encode()).hexdigest())


def train_shapes(cfg, B, stride):
    A = cfg.network.NUM_ANCHORS
    H = 512 // stride
    d = {"data": (B, 3, 512, 512), "im_info": (B, 3), "gt_boxes": (B, 100, 5), "valid_ranges": (B, 2),
         "label": (B, A * H * H), "bbox_target": (B, 4 * A, H, H), "bbox_weight": (B, 4 * A, H, H)}
    return d


def build():
    from sniper_b200 import mxnet_compat as MC
    out = {}
    res = MC.load_symbol_file(os.path.join(REF, "symbols/faster/resnet_mx_101_e2e.py"))
    for tag, fp16 in (("resnet101_train_fp32", False), ("resnet101_train_fp16", True)):
        cfg = load_config("sniper_res101_e2e.yml")
        cfg.TRAIN.fp16 = fp16
        cfg.TRAIN.BATCH_IMAGES = 20
        with MC.NameManager():
            inst = res.resnet_mx_101_e2e(n_proposals=400, momentum=0.995)
            sym = inst.get_symbol_rcnn(cfg)
        out[tag] = describe(sym, train_shapes(cfg, 20, 16))
        out[tag]["cfg"] = dict(fp16=bool(fp16), batch_images=20, num_anchors=int(cfg.network.NUM_ANCHORS),
                               num_classes=int(cfg.dataset.NUM_CLASSES))
    # the AutoFocus inference graph (configs/faster/sniper_res101_e2e_autofocus.yml would set TEST.AUTO_FOCUS)
    cfg = load_config("sniper_res101_e2e.yml")
    cfg.TRAIN.fp16 = False
    cfg.TEST.AUTO_FOCUS = True
    with MC.NameManager():
        inst = res.resnet_mx_101_e2e(n_proposals=400, momentum=0.995, test_nbatch=2)
        sym = inst.get_symbol_rcnn(cfg, is_train=False)
    out["resnet101_test_autofocus"] = describe(sym, {"data": (2, 3, 512, 512), "im_info": (2, 3), "im_ids": (2,), "chip_ids": (2,)})
    # the RPN-only graphs (get_symbol_rpn :157-225)
    cfg = load_config("sniper_res101_e2e.yml")
    cfg.TRAIN.fp16 = False
    cfg.TRAIN.BATCH_IMAGES = 20
    with MC.NameManager():
        sym = res.resnet_mx_101_e2e(n_proposals=400, momentum=0.995).get_symbol_rpn(cfg)
    d = train_shapes(cfg, 20, 16)
    out["resnet101_rpn_train"] = describe(sym, {k: d[k] for k in ("data", "label", "bbox_target", "bbox_weight")})
    with MC.NameManager():
        sym = res.resnet_mx_101_e2e(n_proposals=400, momentum=0.995, test_nbatch=2).get_symbol_rpn(cfg, is_train=False)
    out["resnet101_rpn_test"] = describe(sym, {"data": (2, 3, 512, 512), "im_info": (2, 3), "im_ids": (2,)})
    mob = MC.load_symbol_file(os.path.join(REF, "symbols/faster/mobilenetv2_e2e.py"))
    cfg = load_config("sniper_mobilenetv2_e2e.yml")
    cfg.TRAIN.BATCH_IMAGES = 40
    with MC.NameManager():
        inst = mob.mobilenetv2_e2e(n_proposals=400, momentum=0.995)
        sym = inst.get_symbol_rcnn(cfg)
    d = train_shapes(cfg, 40, int(cfg.network.RPN_FEAT_STRIDE))
    d["crowd_boxes"] = (40, 10, 5)
    out["mobilenetv2_train"] = describe(sym, d)
    out["mobilenetv2_train"]["cfg"] = dict(fp16=bool(cfg.TRAIN.fp16), batch_images=40, num_anchors=int(cfg.network.NUM_ANCHORS),
                                           num_classes=int(cfg.dataset.NUM_CLASSES), feat_stride=int(cfg.network.RPN_FEAT_STRIDE),
                                           scales=list(cfg.network.ANCHOR_SCALES), ratios=list(cfg.network.ANCHOR_RATIOS),
                                           fixed_params=list(cfg.network.FIXED_PARAMS))
    return out


if __name__ == "__main__":
    res = build()
    path = os.path.join(ROOT, "tests", "golden", "ref_symbols.json")
    with open(path, "w") as f:
        json.dump(res, f, sort_keys=True, separators=(",", ":"))      # compact: ~150 KB
    for k, v in res.items():
        print(k, len(v["arguments"]), "args", len(v["auxiliary"]), "aux", v["outputs"], v["ops"])
