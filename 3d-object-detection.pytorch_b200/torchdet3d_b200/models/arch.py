"""Backbone tables: the in-repo MobileNetV3 variants of the reference and torchvision's EfficientNet-B0 / B3.

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

# torchvision EfficientNet (BASELINE configs 3, 5; not part of the reference): rows are torchvision's
# MBConvConfig(expand_ratio, kernel, stride, input_channels, out_channels, num_layers) for width = depth = 1.0;
# `efficientnet_blocks` applies the width / depth multipliers the way torchvision does (adjust_channels / adjust_depth).
EFFICIENTNET_ROWS = ((1, 3, 1, 32, 16, 1), (6, 3, 2, 16, 24, 2), (6, 5, 2, 24, 40, 2), (6, 3, 2, 40, 80, 3),
                     (6, 5, 1, 80, 112, 3), (6, 5, 2, 112, 192, 4), (6, 3, 1, 192, 320, 1))
EFFICIENTNET_MULTS = {"efficientnet_b0": (1.0, 1.0), "efficientnet_b3": (1.2, 1.4)}

ARCH_MOBILENETV3, ARCH_EFFICIENTNET = 0, 1
AVAILABLE_MODELS = set(MODEL_TABLES) | set(EFFICIENTNET_MULTS)
# names the reference accepts but whose definitions live in un-vendored third-party packages
# (timm / efficientnet_lite_pytorch, model_builder.py:4-8,14-17): not provided here.
REFERENCE_ONLY_MODELS = {"mobilenetv3_large_21k", "efficientnet-lite0", "efficientnet-lite1",
                         "efficientnet-lite2"}


def round_channels(v, divisor=8, floor=None):
    """Channel rounding rule of the reference (`_make_divisible`, mobilenetv3.py:54-71)."""
    floor = divisor if floor is None else floor
    r = max(floor, int(v + divisor / 2) // divisor * divisor)
    return r + divisor if r < 0.9 * v else r


def efficientnet_blocks(name):
    import math
    width, depth = EFFICIENTNET_MULTS[name]
    ch = lambda c: round_channels(c * width)       # torchvision MBConvConfig.adjust_channels == _make_divisible(c * w, 8)
    out, stem = [], ch(32)
    for stage, (t, k, s, cin, cout, n) in enumerate(EFFICIENTNET_ROWS, start=1):
        cin, cout = ch(cin), ch(cout)
        for j in range(int(math.ceil(n * depth))):
            i_ch = cin if j == 0 else cout
            out.append(dict(kernel=k, stride=s if j == 0 else 1, in_ch=i_ch, exp_ch=round_channels(i_ch * t), out_ch=cout,
                            use_se=1, se_hidden=max(1, i_ch // 4), use_hs=2, name_stage=stage, name_index=j))
    last = 4 * out[-1]["out_ch"]
    return stem, out, last, last


def arch_of(name):
    return ARCH_EFFICIENTNET if name in EFFICIENTNET_MULTS else ARCH_MOBILENETV3


def blocks_for(name):
    """-> (stem_ch, [dict per block], last_ch, head_ch)."""
    if name in EFFICIENTNET_MULTS:
        return efficientnet_blocks(name)
    spec = MODEL_TABLES[name]
    cin = round_channels(16)
    stem = cin
    out = []
    for k, t, c, se, hs, s in spec["rows"]:
        cout = round_channels(c)
        exp = round_channels(cin * t)
        out.append(dict(kernel=k, stride=s, in_ch=cin, exp_ch=exp, out_ch=cout, use_se=int(bool(se)),
                        se_hidden=round_channels(exp // 4) if se else 0, use_hs=int(bool(hs)), name_stage=0, name_index=0))
        cin = cout
    return stem, out, out[-1]["exp_ch"], spec["head_ch"]
