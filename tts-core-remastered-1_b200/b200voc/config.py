"""GANConfig: host-side mirror of the reference dataclass (vocoder7/config.py:6-40), field for
field, plus the fields the reference reads but never defines (hidden_dim, repair R1) and the
B200 knobs.  A reference ``GANConfig`` instance (or any object with these attributes) is accepted
wherever this class is."""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional


@dataclass
class GANConfig:
    channels: int = 80
    cond_dim: int = 128
    style_dim: int = 128
    num_bands: int = 4
    upsample_factors: List[int] = None
    res_dilations: List[int] = None
    disc_periods: List[int] = None
    disc_kernel_sizes: List[int] = None
    sr: int = 22050
    hop_length: int = 256
    stft_sizes: List[int] = None
    num_style_tokens: int = 10
    dropout_prob: float = 0.1
    r1_gamma: float = 10.0
    r1_interval: int = 16
    lambda_stft: float = 2.0
    lambda_pitch: float = 1.0
    lambda_dur: float = 1.0
    # repairs: read by vocoder7/generator.py:19,31 / implied by generator.py:43-44
    hidden_dim: int = 512
    use_attention: bool = True
    attn_window: Optional[int] = None     # None = global attention over all positions
    # B200 knob: tensor-core operand / activation storage plan ("fp16" | "bf16" | "mixed")
    precision: str = "fp16"

    def __post_init__(self):
        if self.upsample_factors is None:
            self.upsample_factors = [8, 8, 2, 2]
        if self.res_dilations is None:
            self.res_dilations = [1, 3, 5]
        if self.disc_periods is None:
            self.disc_periods = [2, 3, 5, 7, 11]
        if self.disc_kernel_sizes is None:
            self.disc_kernel_sizes = [15, 41, 41]
        if self.stft_sizes is None:
            self.stft_sizes = [512, 1024, 2048]
