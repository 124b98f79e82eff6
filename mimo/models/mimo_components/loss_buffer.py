"""Loss buffer with the reference's interface (reference: mimo/models/mimo_components/loss_buffer.py) whose
state lives on the GPU: no device<->host round trip per training step.

Semantics kept bit-for-bit in spirit: ring buffer of the last `buffer_size` per-subnetwork losses initialised
with zeros, mean over ALL rows (zeros included), weights = softmax(mean / T) * S, `buffer_size == 0` -> ones.
"""
import torch


def softmax_temperature(x: torch.Tensor, temperature=1.0):
    assert temperature > 0, "Temperature should be positive."
    return torch.softmax(x / temperature, dim=-1)


class LossBuffer:
    def __init__(self, subnetworks: int, temperature: float, buffer_size: int, device=None) -> None:
        assert temperature > 0, "Temperature should be positive."
        self.temperature = temperature
        self.buffer_size = buffer_size
        self.subnetworks = subnetworks
        self._device = torch.device(device) if device is not None else None
        self._dev_state = None  # mimo_unet_b200.functional.DeviceLossBuffer, created on first use

    # -- device state ---------------------------------------------------------------------------
    def device_state(self, device=None):
        """The device-resident state (created lazily on the device of the first loss / request)."""
        from mimo_unet_b200.functional import DeviceLossBuffer
        if self._dev_state is None:
            dev = torch.device(device) if device is not None else (self._device or torch.device("cuda"))
            self._dev_state = DeviceLossBuffer(self.subnetworks, self.temperature, self.buffer_size, dev)
        return self._dev_state

    # -- reference interface ----------------------------------------------------------------------
    @property
    def index(self) -> int:
        return 0 if self._dev_state is None else self._dev_state.index

    @property
    def buffer(self) -> torch.Tensor:
        if self._dev_state is None:
            return torch.zeros(self.buffer_size, self.subnetworks)
        return self._dev_state.buffer

    def add(self, loss: torch.Tensor) -> None:
        self.device_state(loss.device if loss.is_cuda else None).add(loss)

    def get_mean(self) -> torch.Tensor:
        if self.buffer_size == 0:
            return torch.zeros(self.subnetworks, device=self.buffer.device)
        return self.buffer.mean(dim=0)

    def get_weights(self) -> torch.Tensor:
        return self.device_state().weights()
