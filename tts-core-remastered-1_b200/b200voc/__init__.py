"""b200voc -- B200-native (sm_100a) implementation of the vocoder7 waveform-synthesis hot path of
ChiefTriston/TTS-Core-Remastered-1: Generator inference and the STFT / mel / iSTFT transforms,
behind the reference's Python API.  See DESIGN.md."""
from .config import GANConfig
from .generator import Generator, ResidualBlock, SelfAttention
from .gst import GlobalStyleTokens
from .discriminators import MultiPeriodDiscriminator, MultiScaleDiscriminator, MultiBandDiscriminator
from .stft import LearnableSTFT, STFTLoss, stft, istft, mel_spectrogram, log_mel, stft_magnitude, stft_prepare
from . import _lib

__all__ = ["GANConfig", "Generator", "GlobalStyleTokens", "MultiPeriodDiscriminator", "MultiScaleDiscriminator",
           "MultiBandDiscriminator", "ResidualBlock", "SelfAttention", "LearnableSTFT", "STFTLoss", "stft", "istft",
           "mel_spectrogram", "log_mel", "stft_magnitude", "stft_prepare"]
