"""GANConfig: host-side mirror of the reference dataclass (vocoder7/config.py:6-40) -- the same field names,
order and defaults, so positional / keyword construction and ``dataclasses.replace`` behave alike -- plus the
fields the reference reads but never defines (hidden_dim, repair R1; use_attention / attn_window, D3) and the B200
knob.  A reference ``GANConfig`` instance (or any object with these attributes) is accepted wherever this class is.

The class is generated from the table below (one row per field: name, type, default, where the reference reads
it); list-valued fields default to ``None`` and are filled in after construction, as in the reference."""
from __future__ import annotations

import dataclasses
from typing import List, Optional

# name, type, default, consumer in the reference
_FIELDS = (
    ("channels", int, 80, "mel bins: generator.py:21, gst.py:19"),
    ("cond_dim", int, 128, "conditioning width: generator.py:26-29"),
    ("style_dim", int, 128, "style vector: generator.py:17, gst.py:17"),
    ("num_bands", int, 4, "band split / merge: generator.py:21-23,51; chunks of the MBD critic"),
    ("upsample_factors", List[int], None, "ConvT strides per stage: generator.py:33-38"),
    ("res_dilations", List[int], None, "dilations of the residual blocks: generator.py:40-41"),
    ("disc_periods", List[int], None, "MPD periods: discriminators.py:16"),
    ("disc_kernel_sizes", List[int], None, "MSD kernel sizes: discriminators.py:71"),
    ("sr", int, 22050, "sampling rate"),
    ("hop_length", int, 256, "= product of upsample_factors"),
    ("stft_sizes", List[int], None, "STFTLoss resolutions: stft.py:42"),
    ("num_style_tokens", int, 10, "gst.py:17"),
    ("dropout_prob", float, 0.1, "classifier-free guidance drop rate (trainer)"),
    ("r1_gamma", float, 10.0, "trainer"),
    ("r1_interval", int, 16, "trainer"),
    ("lambda_stft", float, 2.0, "losses.py"),
    ("lambda_pitch", float, 1.0, "losses.py"),
    ("lambda_dur", float, 1.0, "losses.py"),
    # ---- not in the reference dataclass -------------------------------------------------------------
    ("hidden_dim", int, 512, "read by generator.py:19,31 but never defined (repair R1)"),
    ("use_attention", bool, True, "generator.py:43-44 builds the layer unconditionally (D3)"),
    ("attn_window", Optional[int], None, "None = global attention over all positions (D3)"),
    ("precision", str, "fp16", 'tensor-core operand / activation storage plan: "fp16" | "bf16" | "mixed"'),
)
_LIST_DEFAULTS = {
    "upsample_factors": (8, 8, 2, 2),
    "res_dilations": (1, 3, 5),
    "disc_periods": (2, 3, 5, 7, 11),
    "disc_kernel_sizes": (15, 41, 41),
    "stft_sizes": (512, 1024, 2048),
}


def _fill_list_defaults(self) -> None:
    for name, default in _LIST_DEFAULTS.items():
        if getattr(self, name) is None:
            setattr(self, name, list(default))


GANConfig = dataclasses.make_dataclass(
    "GANConfig",
    [(name, typ, dataclasses.field(default=default)) for name, typ, default, _ in _FIELDS],
    namespace={"__post_init__": _fill_list_defaults, "__doc__": __doc__},
    module=__name__,
)
