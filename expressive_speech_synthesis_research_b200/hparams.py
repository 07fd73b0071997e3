"""Vocoder constants with the reference's names and values (WaveRNN/hparams.py:15-54)."""
sample_rate = 16000
num_mels = 80
hop_length = 200
bits = 9
mu_law = True

voc_mode = 'MOL'
voc_upsample_factors = (5, 5, 8)
voc_rnn_dims = 512
voc_fc_dims = 512
voc_compute_dims = 128
voc_res_out_dims = 128
voc_res_blocks = 10
voc_pad = 2

voc_gen_batched = True
voc_target = 11_000
voc_overlap = 550
