"""Tokenizers: the model directory's HF tokenizer when it has one, otherwise a synthetic byte-level
tokenizer (no tokenizer files exist offline; GPUStack only hands the engine a local HF-format
directory, gpustack/worker/backends/custom.py:91-134).  Also the chat template rendering used by
/v1/chat/completions (the reference engine's job: $VLLM/entrypoints/openai/chat_completion)."""
from __future__ import annotations

from pathlib import Path
from typing import List, Optional, Sequence


class ByteTokenizer:
    """UTF-8 bytes -> ids [OFFSET, OFFSET+256); ids outside that range decode to nothing.
    Deterministic and model-free, so synthetic (random-init) checkpoints can be served."""

    OFFSET = 3  # 0 = pad, 1 = bos, 2 = eos

    def __init__(self, vocab_size: int, bos_token_id: Optional[int] = 1, eos_token_id: Optional[int] = 2):
        if vocab_size < self.OFFSET + 256:
            raise ValueError("vocab too small for the byte tokenizer")
        self.vocab_size = vocab_size
        self.bos_token_id, self.eos_token_id = bos_token_id, eos_token_id

    def encode(self, text: str, add_bos: bool = True) -> List[int]:
        ids = [b + self.OFFSET for b in text.encode("utf-8")]
        return ([self.bos_token_id] if add_bos and self.bos_token_id is not None else []) + ids

    def decode(self, ids: Sequence[int]) -> str:
        lo, hi = self.OFFSET, self.OFFSET + 256
        return bytes(i - lo for i in ids if lo <= i < hi).decode("utf-8", errors="replace")

    def apply_chat_template(self, messages: Sequence[dict]) -> List[int]:
        parts = []
        for m in messages:
            content = m.get("content") or ""
            if isinstance(content, list):  # OpenAI content parts
                content = "".join(p.get("text", "") for p in content if isinstance(p, dict))
            parts.append(f"<|{m.get('role', 'user')}|>\n{content}\n")
        parts.append("<|assistant|>\n")
        return self.encode("".join(parts))


class HFTokenizer:
    def __init__(self, model_path: str):
        from transformers import AutoTokenizer

        self.tok = AutoTokenizer.from_pretrained(model_path)
        self.vocab_size = len(self.tok)
        self.bos_token_id, self.eos_token_id = self.tok.bos_token_id, self.tok.eos_token_id

    def encode(self, text: str, add_bos: bool = True) -> List[int]:
        return self.tok.encode(text, add_special_tokens=add_bos)

    def decode(self, ids: Sequence[int]) -> str:
        return self.tok.decode(list(ids), skip_special_tokens=True)

    def apply_chat_template(self, messages: Sequence[dict]) -> List[int]:
        if getattr(self.tok, "chat_template", None):
            out = self.tok.apply_chat_template(list(messages), add_generation_prompt=True, tokenize=True)
            return list(out["input_ids"] if isinstance(out, dict) else out)
        text = "".join(f"<|{m.get('role', 'user')}|>\n{m.get('content', '')}\n" for m in messages)
        return self.encode(text + "<|assistant|>\n")


def load_tokenizer(model_path: Optional[str], vocab_size: int, bos: Optional[int], eos: Optional[int]):
    if model_path:
        p = Path(model_path)
        if (p / "tokenizer.json").exists() or (p / "tokenizer.model").exists():
            return HFTokenizer(model_path)
    return ByteTokenizer(vocab_size, bos if bos is not None and bos < vocab_size else 1,
                         eos if eos is not None and eos < vocab_size else 2)


class IncrementalDetokenizer:
    """Streams text for one request with a sliding (prefix, read) window so that each push decodes
    only a handful of tokens; holds output back while the tail is an incomplete UTF-8 sequence."""

    def __init__(self, tokenizer):
        self.tok = tokenizer
        self.ids: List[int] = []
        self.prefix = 0
        self.read = 0

    def push(self, new_ids: Sequence[int], final: bool = False) -> str:
        self.ids.extend(new_ids)
        prefix_text = self.tok.decode(self.ids[self.prefix:self.read]) if self.read > self.prefix else ""
        new_text = self.tok.decode(self.ids[self.prefix:])
        if len(new_text) > len(prefix_text) and (final or not new_text.endswith("\ufffd")):
            delta = new_text[len(prefix_text):]
            self.prefix = max(self.read - 4, self.prefix) if self.read - self.prefix > 8 else self.prefix
            self.read = len(self.ids)
            return delta
        return ""
