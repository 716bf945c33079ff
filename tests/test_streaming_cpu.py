"""Host logic of the streaming session (hop scheduling, latent bookkeeping, decode windows) with stand-in flow / decoder
objects on CPU: no CUDA involved.  The stand-ins are causal in the way the real modules are (frame i depends on tokens
<= i/2 + look-ahead; sample s depends on frames within +-3), so the session's output must equal the one-shot result."""
import torch

from minimax_speech_b200.streaming import StreamingSession


class FakeFlow:
    pre_lookahead_len, token_latent_ratio, output_size = 3, 2, 4

    def __init__(self):
        self.calls = []

    def inference(self, token, token_len, prompt_token, prompt_token_len, prompt_feat, prompt_feat_len, embedding=None,
                  reference_mels=None, streaming=False, finalize=False):
        assert streaming and int(token_len[0]) == token.shape[1]
        self.calls.append((token.shape[1], finalize))
        t = token[0, :token.shape[1] - (0 if finalize else self.pre_lookahead_len)].float()
        frames = torch.repeat_interleave(t, 2) + 0.5 * torch.arange(2 * t.shape[0]).remainder(2)
        return frames.reshape(1, 1, -1).repeat(1, self.output_size, 1), None


class FakeDac:
    hop_length = 4

    def decode(self, z):
        x = torch.nn.functional.pad(z[:, :1], (3, 3))  # reach of 3 frames either side
        y = sum(x[:, :, k:k + z.shape[2]] * (k + 1) for k in range(7))
        return torch.repeat_interleave(y, self.hop_length, dim=2) + torch.arange(self.hop_length).repeat(z.shape[2])


def run(n_tokens, n_prompt, piece):
    flow, dac = FakeFlow(), FakeDac()
    tok = torch.arange(100, 100 + n_tokens)
    sess = StreamingSession(flow, dac, torch.zeros(1, n_prompt, dtype=torch.int64), torch.zeros(1, 2 * n_prompt, 4), dac_context=3)
    chunks = []
    for i in range(0, n_tokens, piece):
        chunks += sess.push(tok[i:i + piece].tolist())
    chunks.append(sess.finish())
    return flow, dac, sess, torch.cat(chunks, dim=1), tok


def test_hop_schedule_and_equivalence():
    flow, dac, sess, wav, tok = run(n_tokens=103, n_prompt=10, piece=7)
    # first hop 25 + 15 (prompt padded to the 25-token chunk), then 25, 25, then the finalize call on everything
    assert flow.calls == [(43, False), (68, False), (93, False), (103, True)]
    whole, _ = FakeFlow().inference(tok.unsqueeze(0), torch.tensor([103]), None, None, None, None, streaming=True, finalize=True)
    assert torch.equal(sess.latents, whole)
    assert torch.equal(wav, dac.decode(whole)[:, 0, :])


def test_exact_multiple_and_single_push():
    flow, dac, sess, wav, tok = run(n_tokens=75, n_prompt=25, piece=75)
    # 75 tokens: hops at 25 and 50 need 3 look-ahead tokens each; the third hop has none left -> goes out in finish()
    assert flow.calls == [(28, False), (53, False), (75, True)]
    assert wav.shape == (1, 75 * 2 * 4) and sess.emitted == 150


def test_shorter_than_one_hop():
    flow, dac, sess, wav, tok = run(n_tokens=9, n_prompt=0, piece=4)
    assert flow.calls == [(9, True)] and wav.shape == (1, 9 * 2 * 4)
