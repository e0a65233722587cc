from .utils import (BankLoader, FeatureBank, IndexBatch, TextTensorDataset,  # noqa: F401
                    get_few_shot_setup_name, local_slice, shard_bank)
