"""GANConfig: host-side mirror of the reference dataclass (vocoder7/config.py:6-40) -- the same field names, order
and defaults, so positional / keyword construction and ``dataclasses.replace`` behave alike -- plus the fields the
reference reads but never defines (hidden_dim, repair R1; use_attention / attn_window, D3) and the B200 knob.  A
reference ``GANConfig`` instance (or any object with these attributes) is accepted wherever this class is."""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional


@dataclass
class GANConfig:
    channels: int = 80                            # mel bins: generator.py:21, gst.py:19
    cond_dim: int = 128                           # conditioning width: generator.py:26-29
    style_dim: int = 128                          # style vector: generator.py:17, gst.py:17
    num_bands: int = 4                            # band split / merge: generator.py:21-23,51; chunks of the MBD critic
    upsample_factors: Optional[List[int]] = None  # ConvT strides per stage: generator.py:33-38
    res_dilations: Optional[List[int]] = None     # dilations of the residual blocks: generator.py:40-41
    disc_periods: Optional[List[int]] = None      # MPD periods: discriminators.py:16
    disc_kernel_sizes: Optional[List[int]] = None # MSD kernel sizes: discriminators.py:71
    sr: int = 22050
    hop_length: int = 256                         # = product of upsample_factors
    stft_sizes: Optional[List[int]] = None        # STFTLoss resolutions: stft.py:42
    num_style_tokens: int = 10                    # gst.py:17
    dropout_prob: float = 0.1                     # classifier-free guidance drop rate (trainer)
    r1_gamma: float = 10.0
    r1_interval: int = 16
    lambda_stft: float = 2.0
    lambda_pitch: float = 1.0
    lambda_dur: float = 1.0
    # ---- not in the reference dataclass
    hidden_dim: int = 512                         # read by generator.py:19,31 but never defined (repair R1)
    use_attention: bool = True                    # generator.py:43-44 builds the layer unconditionally (D3)
    attn_window: Optional[int] = None             # None = global attention over all positions (D3)
    precision: str = "fp16"                       # tensor-core operand / activation storage plan: "fp16" | "bf16" | "mixed"

    def __post_init__(self):
        # list-valued fields are filled in after construction, as in the reference (config.py:30-40)
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
        self.validate()

    # ---- what the kernels need from a configuration (checked once here, loudly, instead of deep inside a launch)
    def validate(self) -> None:
        if self.channels % self.num_bands != 0:
            raise ValueError(f"channels={self.channels} must be a multiple of num_bands={self.num_bands}")
        if self.hidden_dim % (1 << len(self.upsample_factors)) != 0:
            raise ValueError(f"hidden_dim={self.hidden_dim} is halved {len(self.upsample_factors)} times (generator.py:36,46)")
        if self.precision not in ("fp16", "bf16", "mixed"):
            raise ValueError(f"unknown precision plan {self.precision!r}")
        if self.attn_window is not None and self.attn_window <= 0:
            raise ValueError("attn_window must be positive (or None for global attention)")
        if any(f < 2 or f % 2 for f in self.upsample_factors):
            raise ValueError(f"upsample_factors {self.upsample_factors}: every stride must be even (k = 2f, p = f/2)")

    @property
    def hop(self) -> int:
        """samples per mel frame = product of the upsampling strides (the reference keeps a separate hop_length)"""
        h = 1
        for f in self.upsample_factors:
            h *= f
        return h

    @classmethod
    def from_reference(cls, ref, **overrides) -> "GANConfig":
        """Build from a reference ``vocoder7.config.GANConfig`` instance (or any object with those attributes); the
        fields the reference lacks keep their defaults unless overridden."""
        import dataclasses
        kw = {f.name: getattr(ref, f.name) for f in dataclasses.fields(cls) if hasattr(ref, f.name)}
        kw.update(overrides)
        return cls(**kw)
