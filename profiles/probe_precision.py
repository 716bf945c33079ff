import sys, torch, os
sys.path.insert(0, "/root/repo")
import minimax_speech_b200.synth as synth
from minimax_speech_b200.flow import ConditionalCFM, CausalConditionalDecoder
from oracle import restatement as O
torch.set_num_threads(os.cpu_count())
DEV = torch.device("cuda:0")
sd = synth.estimator_state_dict(7, init="test")
est = CausalConditionalDecoder(); est.load_state_dict(sd)
cfm = ConditionalCFM(240, dict(t_scheduler="cosine", inference_cfg_rate=0.7), 1, 80, est)
for T in (70, 150, 300):
    for temp in (0.8, 1.0):
        for idx in (71, 5):
            mu, mask, spks, cond = synth.batch_inputs([T], first_index=idx)
            g = torch.Generator().manual_seed(500 + idx); z = torch.randn(1, 80, T, generator=g)
            y, _ = cfm(mu.clone().to(DEV), mask.to(DEV), 10, temperature=temp, spks=spks.to(DEV), cond=cond.to(DEV), noise=z)
            with torch.inference_mode():
                ref, _ = O.cfm_forward_cached(sd, z, mu, mask, 10, temp, spks, cond)
            print(f"T={T} temp={temp} idx={idx}: rel-L2 {O.rel_l2(y.cpu(), ref):.3e}  |y| {float(ref.abs().mean()):.3f}")
