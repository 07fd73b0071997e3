"""`Synthesize` facade with the reference's shape (WaveRNN/synthesizer_wavernn.py:8-33), so the
sentence drivers (synthesize_sentences.py:47-73) switch by changing one import."""
import torch

from . import hparams as hp
from .wavernn import WaveRNN


class Synthesize:

    def __init__(self, model_path=None, hparams=hp, device=None):
        if device is None:
            if not torch.cuda.is_available():
                raise RuntimeError("the B200 WaveRNN vocoder needs a CUDA device (no CPU path)")
            device = torch.device('cuda')
        self.hp = hparams
        self.voc_model = WaveRNN(rnn_dims=hparams.voc_rnn_dims,
                                 fc_dims=hparams.voc_fc_dims,
                                 bits=hparams.bits,
                                 pad=hparams.voc_pad,
                                 upsample_factors=hparams.voc_upsample_factors,
                                 feat_dims=hparams.num_mels,
                                 compute_dims=hparams.voc_compute_dims,
                                 res_out_dims=hparams.voc_res_out_dims,
                                 res_blocks=hparams.voc_res_blocks,
                                 hop_length=hparams.hop_length,
                                 sample_rate=hparams.sample_rate,
                                 mode=hparams.voc_mode).to(device)
        if model_path is not None:
            self.voc_model.restore(model_path)

    def generate(self, mel, batch_pred=True):
        # reads target / overlap / mu_law at call time like the reference (:32)
        return self.voc_model.generate(mel, batch_pred, self.hp.voc_target, self.hp.voc_overlap, self.hp.mu_law)

    def generate_many(self, mels):
        """Pooled-fold synthesis of a sentence set (one launch sequence for all utterances)."""
        return self.voc_model.generate_many(mels, self.hp.voc_target, self.hp.voc_overlap, self.hp.mu_law)
