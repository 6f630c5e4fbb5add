"""N-best list for beam search (parity: text_generation/beam_utils.py:19-64; scores are length-normalised sums of
log-probs)."""
from __future__ import annotations


class BeamHypotheses:
    def __init__(self, num_beams, length_penalty=1.0, early_stopping=False):
        self.length_penalty, self.early_stopping, self.num_beams = length_penalty, early_stopping, num_beams
        self.beams = []
        self.worst_score = 1e9

    def __len__(self):
        return len(self.beams)

    def add(self, hyp, sum_logprobs, length):
        score = sum_logprobs / length ** self.length_penalty
        if len(self) < self.num_beams or score > self.worst_score:
            self.beams.append((score, hyp))
            if len(self) > self.num_beams:
                order = sorted((s, i) for i, (s, _) in enumerate(self.beams))
                del self.beams[order[0][1]]
                self.worst_score = order[1][0]
            else:
                self.worst_score = min(score, self.worst_score)

    def is_done(self, best_sum_logprobs, cur_len):
        """True when no running hypothesis can still beat the worst kept one."""
        if len(self) < self.num_beams:
            return False
        if self.early_stopping:
            return True
        return self.worst_score >= best_sum_logprobs / cur_len ** self.length_penalty
