"""Deterministic stand-in for the Phi-3.5-vision tokenizer (its files are not available offline): the chat template
string of one user turn, an eos token, and a hashing word tokenizer. Only the *structure* of the prompt matters to the
scoring path (bos, <|user|>, image slots, caption ids, eos); the caption ids are arbitrary valid vocabulary ids."""
import re
import types
import zlib


class StubPhi3Tokenizer:
    bos_token_id, pad_token_id, eos_token_id = 1, 32000, 32000
    eos_token, pad_token = "<|endoftext|>", "<|endoftext|>"
    padding_side, truncation_side = "left", "right"
    SPECIAL = {"<|user|>": 32010, "<|end|>": 32007, "<|assistant|>": 32001, "<|endoftext|>": 32000}

    def apply_chat_template(self, messages, tokenize=False, add_generation_prompt=True, **kw):
        assert not tokenize and len(messages) == 1 and messages[0]["role"] == "user"
        text = f"<|user|>\n{messages[0]['content']}<|end|>\n"
        return text + ("<|assistant|>\n" if add_generation_prompt else "")

    def __call__(self, text, **kw):
        ids = []
        for piece in re.findall(r"<\|[a-z_0-9]+\|>|\n|[^\s<]+|<", text):
            if piece in self.SPECIAL:
                ids.append(self.SPECIAL[piece])
            elif piece == "\n":
                ids.append(13)
            else:
                ids.append(3 + zlib.crc32(piece.encode()) % 31000)
        # like the Llama tokenizer the reference uses, every chunk starts with bos (processing_phi3_v.py:425 keeps it)
        return types.SimpleNamespace(input_ids=[self.bos_token_id] + ids)
