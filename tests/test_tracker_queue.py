"""usot_b200.tracker_ops.MemoryQueue reproduces the reference tracker's memory-queue sampling (lib/tracker/usot_tracker.py:222-256),
restated here with host lists exactly as the reference keeps them."""
import numpy as np
import torch

from usot_b200.tracker_ops import MemoryQueue


def _reference_select(init_features, memory_features, memory_confidences, mem_queue_size=7):
    template_mem = list(init_features)
    score_mem = [0.9, 0.9]
    mem_length = len(memory_confidences)
    upd = mem_queue_size - 3
    if mem_length <= 1:
        template_mem += [memory_features[0]] * (upd + 1)
        score_mem += [memory_confidences[0]] * (upd + 1)
    else:
        gap = (mem_length - 1) / upd
        for i in range(upd):
            start_index = min(int(int(i * gap) * mem_length), mem_length - 1)
            end_index = min(int(int((i + 1) * gap) * mem_length), mem_length - 1)
            if start_index >= end_index:
                template_mem.append(memory_features[start_index])
                score_mem.append(memory_confidences[start_index])
            else:
                score_tmp = np.array(memory_confidences[start_index:end_index])
                max_index = np.argmax(score_tmp) + start_index
                template_mem.append(memory_features[max_index])
                score_mem.append(memory_confidences[max_index])
        template_mem.append(memory_features[-1])
        score_mem.append(memory_confidences[-1])
    return torch.cat(template_mem, dim=0), torch.tensor(score_mem).unsqueeze(0)


def test_memory_queue_matches_reference_sampling():
    g = torch.Generator().manual_seed(0)
    feat = lambda: torch.randn(1, 256, 7, 7, generator=g)
    f0, f0_flip = feat(), feat()
    q = MemoryQueue([f0, f0_flip], mem_queue_size=7, capacity=4)  # small capacity: exercises the growth path
    init, mems, confs = [f0, f0_flip], [f0], [0.9]
    rng = np.random.RandomState(1)
    for frame in range(40):
        mem, score = q.select()
        ref_mem, ref_score = _reference_select(init, mems, confs)
        assert tuple(mem.shape) == (7, 256, 7, 7) and tuple(score.shape) == (1, 7)
        assert torch.equal(mem.contiguous(), ref_mem), f"frame {frame}"
        assert torch.allclose(score, ref_score)
        f, c = feat(), float(rng.rand())
        q.append(f, c)
        mems.append(f)
        confs.append(c)
