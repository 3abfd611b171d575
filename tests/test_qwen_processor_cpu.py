"""Host logic of the Qwen2.5-VL processor (no GPU): <|image_pad|> expansion vs the merged-token count the model
expects, the chat-template fallback vs the slice the reference applies (reward_dataset.py:417)."""
import pytest
import torch

from llava_reward_b200.processing import Qwen2_5_VLProcessorB200, smart_resize


class FakeTok:
    pad_token_id, padding_side, chat_template = 151643, "left", None

    def __call__(self, texts, padding=False, return_tensors="pt", **kw):
        rows = []
        for t in texts:
            ids, i = [], 0
            while i < len(t):
                if t.startswith("<|image_pad|>", i):
                    ids.append(151655)
                    i += 13
                else:
                    ids.append(10 + ord(t[i]) % 50)
                    i += 1
            rows.append(ids)
        S = max(len(r) for r in rows)
        ids = torch.tensor([[151643] * (S - len(r)) + r for r in rows])
        mask = torch.tensor([[0] * (S - len(r)) + [1] * len(r) for r in rows])
        return {"input_ids": ids, "attention_mask": mask}


class StubImageProc:
    merge_size = 2

    def __call__(self, images, return_tensors="pt"):
        grids = []
        for im in images:
            rh, rw = smart_resize(im.shape[0], im.shape[1])
            grids.append([1, rh // 14, rw // 14])
        return {"pixel_values": torch.zeros(sum(g[1] * g[2] for g in grids), 4), "image_grid_thw": torch.tensor(grids)}


def test_image_pad_expansion_and_batch_layout():
    proc = Qwen2_5_VLProcessorB200(StubImageProc(), FakeTok())
    imgs = [torch.zeros(512, 640, 3), torch.zeros(200, 333, 3)]
    texts = ["<|vision_start|><|image_pad|><|vision_end|>a cat", "<|vision_start|><|image_pad|><|vision_end|>dog"]
    out = proc(text=texts, images=imgs, padding=True, return_tensors="pt")
    n = (out["input_ids"] == 151655).sum(1).tolist()
    grids = out["image_grid_thw"].tolist()
    assert n == [g[1] * g[2] // 4 for g in grids] == [36 * 46 // 4, 26 * 42 // 4]
    assert out["attention_mask"][1, 0] == 0 and out["input_ids"][1, 0] == 151643        # left padding
    with pytest.raises(ValueError, match="more"):
        proc(text=[texts[0] + "<|image_pad|>", texts[1]], images=imgs, padding=True)
    with pytest.raises(ValueError, match="fewer"):
        proc(text=[texts[0], "no image"], images=imgs, padding=True)


def test_chat_template_fallback_matches_reference_slice():
    """the reference keeps template[58:-23].strip(): the user turn without the system prompt / generation prompt"""
    proc = Qwen2_5_VLProcessorB200(StubImageProc(), FakeTok())
    msg = [{"role": "user", "content": [{"type": "image", "image": "file://x.jpg"}, {"type": "text", "text": "a red car"}]}]
    t = proc.apply_chat_template(msg, tokenize=False, add_generation_prompt=True)
    assert t[58:-23].strip() == "<|im_start|>user\n<|vision_start|><|image_pad|><|vision_end|>a red car<|im_end|>"
