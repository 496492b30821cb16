"""Birds_Eye_View_Loss/Networks/ERFNet.py mirror: identical blocks; ``Net.forward`` returns
(encoder_output, decoder_output) -- a 2-tuple (reference :151-157) instead of BP's 3-tuple."""
from ._pkg import bp

_E = bp("ERFNet")
DownsamplerBlock = _E.DownsamplerBlock
non_bottleneck_1d = _E.non_bottleneck_1d
Encoder = _E.Encoder
UpsamplerBlock = _E.UpsamplerBlock
Decoder = _E.Decoder


class Net(_E.Net):
    def forward(self, input, flag, only_encode=False):
        if only_encode:
            return self.encoder.forward(input, predict=True)
        encoder_output, decoder_output, _ = super().forward(input, flag)
        return encoder_output, decoder_output
