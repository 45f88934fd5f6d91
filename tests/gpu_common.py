"""Helpers for the `-m gpu` tests: build the drop-in model on cuda:0 from an oracle weight dict."""
import torch

from oracle import vsr_oracle as O


def make_model(d: O.Dims, W, verb_table=None, device="cuda:0"):
    from models import ControllableCaptioningModel
    m = ControllableCaptioningModel(d.seq_len, d.vocab_size, d.bos_idx, det_feat_size=d.det_feat_size,
                                    input_encoding_size=d.input_encoding_size, rnn_size=d.rnn_size,
                                    att_size=d.att_size, h2_first_lstm=d.h2_first_lstm,
                                    img_second_lstm=d.img_second_lstm,
                                    verb_tables=(verb_table or {}, {}))
    missing = m.load_state_dict(W, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    return m.to(device).eval()


def device_beam(m, statics, eos, k, out_size=1, use_verbs=False, gt=False, trace=True):
    """Run the device beam search through the engine, returning outputs, log-probs, history and
    (optionally) per-step log-prob traces."""
    eng = m._engine_for(statics)
    (w, g), (lw, lg), extra = eng.beam_search(k, out_size, eos, use_verbs=use_verbs, gt=gt, trace_steps=trace)
    hist = eng.history()
    torch.cuda.synchronize()
    return (w, g), (lw, lg), hist, extra
