"""Random-init stand-in for the CLAP text query encoder — OFF the hot path, plain PyTorch.

north_star: "The CLAP text encoder remains PyTorch, since it is off the hot path."  The reference's ``CLAP_Encoder``
(``models/clap_encoder.py:10-116``) needs the LAION-CLAP checkpoint, the ``roberta-base`` tokenizer files and ``h5py``,
none of which exist offline, so BASELINE config 3 ("full AudioSep inference, random-init CLAP text encoder") uses this
stand-in with the same architecture and call contract:

* text tower = ``RobertaModel`` (roberta-base geometry: 12 layers, 768 hidden, 12 heads, vocab 50265, 514 positions),
  pooler output -> ``Linear(768, 512) - ReLU - Linear(512, 512)`` -> L2 normalise
  (reference ``models/CLAP/open_clip/model.py:516-531,658-665,732-752``);
* captions are padded to 512 tokens with pad id 1 like the reference (``models/clap_encoder.py:108-116``); because the
  BPE vocabulary files are unavailable, token ids are a deterministic hash of the whitespace-split words (synthetic ids —
  the weights are random anyway);
* ``get_query_embed(modality='text', text=[...])`` returns ``(B, 512)`` float32; embeddings are cached per caption
  (the reference recomputes a 512-token RoBERTa pass per clip).
"""
import zlib
from typing import Dict, List

import torch
import torch.nn as nn
import torch.nn.functional as F


def synthetic_token_ids(text: str, max_length: int = 512, vocab_size: int = 50265) -> torch.Tensor:
    ids = [0]                                                     # <s>
    for word in text.lower().split():
        ids.append(3 + zlib.crc32(word.encode("utf-8")) % (vocab_size - 4))
    ids = ids[: max_length - 1] + [2]                             # </s>
    ids = ids + [1] * (max_length - len(ids))                     # <pad> = 1
    return torch.tensor(ids, dtype=torch.long)


class RandomInitCLAPTextEncoder(nn.Module):
    def __init__(self, hidden_size: int = 768, num_hidden_layers: int = 12, num_attention_heads: int = 12,
                 intermediate_size: int = 3072, joint_embed_shape: int = 512, max_length: int = 512, seed: int = 0):
        super().__init__()
        from transformers import RobertaConfig, RobertaModel
        cfg = RobertaConfig(vocab_size=50265, hidden_size=hidden_size, num_hidden_layers=num_hidden_layers,
                            num_attention_heads=num_attention_heads, intermediate_size=intermediate_size,
                            max_position_embeddings=max_length + 2, type_vocab_size=1, layer_norm_eps=1e-5,
                            pad_token_id=1, bos_token_id=0, eos_token_id=2)
        with torch.random.fork_rng():
            torch.manual_seed(seed)
            self.text_branch = RobertaModel(cfg)
            self.text_projection = nn.Sequential(nn.Linear(hidden_size, joint_embed_shape), nn.ReLU(),
                                                 nn.Linear(joint_embed_shape, joint_embed_shape))
        self.max_length = max_length
        self.encoder_type = "CLAP"
        self._cache: Dict[str, torch.Tensor] = {}
        for p in self.parameters():
            p.requires_grad = False
        self.eval()

    def tokenizer(self, text: List[str]):
        ids = torch.stack([synthetic_token_ids(t, self.max_length) for t in text])
        return {"input_ids": ids, "attention_mask": (ids != 1).long()}

    @torch.no_grad()
    def _get_text_embed(self, batch: List[str]) -> torch.Tensor:
        device = next(self.parameters()).device
        missing = [t for t in dict.fromkeys(batch) if t not in self._cache]
        if missing:
            tok = {k: v.to(device) for k, v in self.tokenizer(missing).items()}
            pooled = self.text_branch(**tok).pooler_output
            emb = F.normalize(self.text_projection(pooled), dim=-1)
            for t, e in zip(missing, emb):
                self._cache[t] = e.detach()
        return torch.stack([self._cache[t].to(device) for t in batch])

    def get_query_embed(self, modality, audio=None, text=None, use_text_ratio=0.5, device=None):
        if modality not in ("text", "hybird"):
            raise NotImplementedError("the stand-in only has the text tower (reference use_text_ratio is 1.0)")
        return self._get_text_embed(list(text)).float()
