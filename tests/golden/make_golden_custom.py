"""Golden fixture for the RMSNorm / SwiGLU / RoPE decoder of VyomAI/models/custom_transformer.py (`ModelForCausalLM`), produced by
running the REAL reference class (imported from /root/reference) on the CPU in fp32:

    python tests/golden/make_golden_custom.py

Two tiny configs: head_dim 64 (hidden 128, 2 q heads / 1 kv head) and head_dim 128 (hidden 256, 2 / 1). Stored: logits of a
right-padded batch (causal x key-padding mask, `use_cache=False`), and greedy ids of an unpadded prompt obtained by repeated
full forwards (no cache on the reference side: its cached path goes through transformers' DynamicCache, whose interface has
moved since the file was written) with the reference's own top-1 / top-2 margins. Weights rounded to bf16-representable values."""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import REF, round_weights_, save  # noqa: E402


def main():
    sys.path.insert(0, REF)
    from VyomAI.models.custom_transformer import Config, ModelForCausalLM
    for name, hidden, heads, kv in (("custom_lm_d64", 128, 2, 1), ("custom_lm_d128", 256, 2, 1)):
        torch.manual_seed(99)
        fields = dict(vocab_size=300, hidden_size=hidden, intermediate_size=2 * hidden, num_hidden_layers=2, num_attention_heads=heads,
                      num_key_value_heads=kv, max_position_embeddings=64, rms_norm_eps=1e-6, rope_theta=1000000.0, tie_word_embeddings=True,
                      pad_token_id=0)
        model = ModelForCausalLM(Config(**fields)).eval()
        with torch.no_grad():
            for n, p in model.named_parameters():
                if n.endswith("layernorm.weight") or n.endswith("model.norm.weight"):
                    p.add_(0.1 * torch.randn_like(p))
        round_weights_(model)
        g = torch.Generator().manual_seed(7)
        B, S = 3, 12
        ids = torch.randint(3, 300, (B, S), generator=g)
        lens = torch.tensor([12, 5, 9])
        mask = (torch.arange(S)[None, :] < lens[:, None]).long()
        ids = torch.where(mask.bool(), ids, torch.zeros_like(ids))
        with torch.no_grad():
            logits = model(input_ids=ids, attention_mask=mask, use_cache=False).logits
            prompt = ids[:1, :6]
            cur, margins = prompt, []
            for _ in range(6):
                lg = model(input_ids=cur, attention_mask=torch.ones_like(cur), use_cache=False).logits[:, -1]
                t2 = lg.topk(2, -1).values
                margins.append(float((t2[:, 0] - t2[:, 1]).min()))
                cur = torch.cat([cur, lg.argmax(-1, keepdim=True)], dim=1)
        blob = {}
        # ModelForCausalLM both INHERITS BaseModel and owns `self.model = BaseModel(config)`; forward only uses the latter, so the
        # inherited `embed_tokens.* / layers.* / norm.*` entries of its state_dict are dead weights and are not stored
        for k, v in model.state_dict().items():
            if k.startswith("model.") or k.startswith("lm_head."):
                blob["w::" + k] = v.detach().bfloat16().view(torch.int16).numpy().view(np.uint16)
        blob["in::input_ids"], blob["in::attention_mask"], blob["in::prompt"] = ids.numpy(), mask.numpy(), prompt.numpy()
        blob["out::logits"], blob["out::generate"] = logits.numpy(), cur.numpy()
        meta = dict(fields, head_dim=hidden // heads, generate_margins=margins, layer_norm_eps=1e-6, hidden_act="silu")
        blob["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
        save(name, blob)


if __name__ == "__main__":
    main()
