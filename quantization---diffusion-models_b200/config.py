"""AwqConfig: field names and defaults of the reference's models/_config.py:8-33, kept verbatim so the
`quant_config` dict a user passes to quantize() means the same thing."""
from dataclasses import asdict, dataclass, field
from typing import Dict, List, Optional


@dataclass
class AwqConfig:
    quant_method: str = field(default="awq")
    zero_point: bool = field(default=True)
    q_group_size: int = field(default=128)
    w_bit: int = field(default=4)
    wv_bit: int = field(default=4)
    a_bit: int = field(default=16)
    version: str = field(default="fake_act")
    modules_to_not_convert: Optional[List] = None
    weight_quant_conv_type: str = field(default="per_channel")
    weight_quant_type: str = field(default="group")
    act_quant_conv_type: str = field(default="per_channel")
    act_quant_conv_group_size: int = field(default=1)
    quantize_act: bool = field(default=False)
    config_file_name = "config.json"

    @classmethod
    def from_dict(cls, quant_config: Dict = {}):
        """_config.py:26-33 (unknown keys raise TypeError exactly like the dataclass constructor there)."""
        if not quant_config:
            return cls()
        cfg = cls(**quant_config)
        cfg.version = cfg.version.lower()
        return cfg

    def to_dict(self):
        return asdict(self)

    def to_transformers_dict(self):
        """_config.py:97-107."""
        return {
            "quant_method": self.quant_method,
            "zero_point": self.zero_point,
            "group_size": self.q_group_size,
            "bits": self.w_bit,
            "version": self.version.lower(),
            "modules_to_not_convert": self.modules_to_not_convert,
        }
