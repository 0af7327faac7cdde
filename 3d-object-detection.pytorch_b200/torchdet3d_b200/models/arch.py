"""Backbone tables of the in-repo MobileNetV3 variants.

Rows are (kernel, expansion t, out channels c, SE, h-swish, stride) exactly as the reference
lists them (torchdet3d/models/mobilenetv3.py:20-52); `blocks_for` expands them the way
MobileNetV3.__init__ does (:176-186) into the flat descriptors the C plan consumes.
"""

MODEL_TABLES = {
    "mobilenetv3_large": dict(head_ch=1280, rows=(
        (3, 1, 16, 0, 0, 1), (3, 4, 24, 0, 0, 2), (3, 3, 24, 0, 0, 1), (5, 3, 40, 1, 0, 2),
        (5, 3, 40, 1, 0, 1), (5, 3, 40, 1, 0, 1), (3, 6, 80, 0, 1, 2), (3, 2.5, 80, 0, 1, 1),
        (3, 2.3, 80, 0, 1, 1), (3, 2.3, 80, 0, 1, 1), (3, 6, 112, 1, 1, 1), (3, 6, 112, 1, 1, 1),
        (5, 6, 160, 1, 1, 2), (5, 6, 160, 1, 1, 1), (5, 6, 160, 1, 1, 1))),
    "mobilenetv3_small": dict(head_ch=1024, rows=(
        (3, 1, 16, 1, 0, 2), (3, 4.5, 24, 0, 0, 2), (3, 3.67, 24, 0, 0, 1), (5, 4, 40, 1, 1, 2),
        (5, 6, 40, 1, 1, 1), (5, 6, 40, 1, 1, 1), (5, 3, 48, 1, 1, 1), (5, 3, 48, 1, 1, 1),
        (5, 6, 96, 1, 1, 2), (5, 6, 96, 1, 1, 1), (5, 6, 96, 1, 1, 1))),
}

AVAILABLE_MODELS = set(MODEL_TABLES)
# names the reference accepts but whose definitions live in un-vendored third-party packages
# (timm / efficientnet_lite_pytorch, model_builder.py:4-8,14-17): not provided here.
REFERENCE_ONLY_MODELS = {"mobilenetv3_large_21k", "efficientnet-lite0", "efficientnet-lite1",
                         "efficientnet-lite2"}


def round_channels(v, divisor=8, floor=None):
    """Channel rounding rule of the reference (`_make_divisible`, mobilenetv3.py:54-71)."""
    floor = divisor if floor is None else floor
    r = max(floor, int(v + divisor / 2) // divisor * divisor)
    return r + divisor if r < 0.9 * v else r


def blocks_for(name):
    """-> (stem_ch, [dict per block], last_ch, head_ch)."""
    spec = MODEL_TABLES[name]
    cin = round_channels(16)
    stem = cin
    out = []
    for k, t, c, se, hs, s in spec["rows"]:
        cout = round_channels(c)
        exp = round_channels(cin * t)
        out.append(dict(kernel=k, stride=s, in_ch=cin, exp_ch=exp, out_ch=cout, use_se=int(bool(se)),
                        se_hidden=round_channels(exp // 4) if se else 0, use_hs=int(bool(hs))))
        cin = cout
    return stem, out, out[-1]["exp_ch"], spec["head_ch"]
