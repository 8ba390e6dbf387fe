"""[T, N, ...] rollout storage, drop-in for the reference ``ExperienceBuffer``
(experience_buffer.py:4-113): same method names, same permutation-walking minibatch sampler
(a device ``randperm`` consumed M entries at a time, re-drawn on wrap, two draws at construction).

The per-key gathers of ``sample`` exist for API compatibility; the training step only asks for the index
slice (``sample_indices``) and gathers the eight keys it needs inside one kernel (csrc/mlp.cu).
"""
import torch


class ExperienceBuffer:
    def __init__(self, buffer_length, batch_size, device, randperm_fn=None):
        self._buffer_length = buffer_length
        self._batch_size = batch_size
        self._device = device
        self._buffer_head = 0
        self._total_samples = 0
        self._buffers = dict()
        self._flat_buffers = dict()
        self._randperm = randperm_fn or (lambda n: torch.randperm(n, device=self._device, dtype=torch.long))
        self._sample_buf = self._randperm(buffer_length * batch_size)
        self._sample_buf_head = 0
        self._reset_sample_buf()

    def add_buffer(self, name, buffer):
        assert len(buffer.shape) >= 2
        assert buffer.shape[0] == self._buffer_length
        assert buffer.shape[1] == self._batch_size
        assert name not in self._buffers
        self._buffers[name] = buffer
        self._flat_buffers[name] = buffer.view([buffer.shape[0] * buffer.shape[1]] + list(buffer.shape[2:]))

    def reset(self):
        self._buffer_head = 0
        self._reset_sample_buf()

    def clear(self):
        self.reset()
        self._total_samples = 0

    def inc(self):
        self._buffer_head = (self._buffer_head + 1) % self._buffer_length
        self._total_samples += self._batch_size

    def get_total_samples(self):
        return self._total_samples

    def get_sample_count(self):
        return min(self._total_samples, self._buffer_length * self._batch_size)

    def get_buffer_head(self):
        return self._buffer_head

    def record(self, name, data):
        assert data.shape[0] == self._batch_size
        self._buffers[name][self._buffer_head] = data

    def get_data(self, name):
        return self._buffers[name]

    def get_data_flat(self, name):
        return self._flat_buffers[name]

    def set_data(self, name, data):
        buf = self.get_data(name)
        assert buf.shape[0] == data.shape[0] and buf.shape[1] == data.shape[1]
        buf[:] = data

    def set_data_flat(self, name, data):
        buf = self.get_data_flat(name)
        assert buf.shape[0] == data.shape[0]
        buf[:] = data

    def sample(self, n):
        idx = self.sample_indices(n)
        return {k: v[idx] for k, v in self._flat_buffers.items()}

    def sample_indices(self, n):
        """Next n entries of the permutation (experience_buffer.py:90-113), as a contiguous int64 tensor."""
        buffer_len = self._sample_buf.shape[0]
        assert n <= buffer_len
        if self._sample_buf_head + n <= buffer_len:
            rand_idx = self._sample_buf[self._sample_buf_head:self._sample_buf_head + n]
            self._sample_buf_head += n
        else:
            # a VIEW, as in the reference: the in-place re-draw below also replaces these tail entries
            head0 = self._sample_buf[self._sample_buf_head:]
            remainder = n - (buffer_len - self._sample_buf_head)
            self._reset_sample_buf()
            rand_idx = torch.cat([head0, self._sample_buf[:remainder]], dim=0)
            self._sample_buf_head = remainder
        return torch.remainder(rand_idx, self.get_sample_count()).contiguous()

    def _reset_sample_buf(self):
        self._sample_buf[:] = self._randperm(self._buffer_length * self._batch_size)
        self._sample_buf_head = 0
