"""`EncoderBase64` of rlsolver/methods/util_evaluator.py:22-65: solution vector <-> base-64 string
(MSB-first big integer over the alphabet 0-9A-Za-z_$), the format of the reference's logs and of
its best-known Gset solutions X_G14 ... X_G70 (util_evaluator.py:258-289).  Host-side, pure
Python integers; kept call-compatible (same wrapping at 120 characters, same zero fill)."""
from __future__ import annotations

from typing import Union

import numpy as np
import torch as th

TEN = th.Tensor
ARY = np.ndarray


class EncoderBase64:
    def __init__(self, encode_len: int):
        num_power = 6
        self.encode_len = encode_len
        self.string_len = -(-encode_len // num_power)      # ceil(encode_len / 6)
        self.base_digits = "0123456789ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz_$"
        self.base_num = len(self.base_digits)
        assert self.base_num == 2 ** num_power
        self._index = {ch: i for i, ch in enumerate(self.base_digits)}

    def bool_to_str(self, x_bool: Union[TEN, ARY]) -> str:
        bits = np.asarray(x_bool.detach().cpu().numpy() if isinstance(x_bool, th.Tensor) else x_bool).astype(bool)
        x_int = int.from_bytes(np.packbits(bits[::-1], bitorder="little").tobytes(), "little") if bits.size else 0
        digits = []
        while True:
            x_int, rem = divmod(x_int, self.base_num)
            digits.append(self.base_digits[rem])
            if x_int == 0:
                break
        x_str = "".join(reversed(digits))
        if len(x_str) > 120:
            x_str = "\n".join(x_str[i:i + 120] for i in range(0, len(x_str), 120))
        if len(x_str) > 64:
            x_str = f"\n{x_str}"
        return x_str.zfill(self.string_len)

    def str_to_bool(self, x_str: str) -> TEN:
        x_b64 = x_str.replace("\n", "").replace(" ", "")
        x_int = 0
        for ch in x_b64:
            x_int = x_int * self.base_num + self._index[ch]
        x_bool = th.zeros(self.encode_len, dtype=th.bool)
        if x_int:
            x_bin = bin(x_int)[2:]
            x_bool[-len(x_bin):] = th.tensor([c == "1" for c in x_bin], dtype=th.bool)
        return x_bool
