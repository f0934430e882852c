"""Configuration dataclasses with the reference's names and fields
(/root/reference/neural_lam/config.py:28-132), minus the YAML/JSON wizardry
(no compute; the models only read `config.training.state_feature_weighting`).
Objects of the reference's own classes are accepted wherever these are."""
import dataclasses
from typing import Dict, Union


@dataclasses.dataclass
class DatastoreSelection:
    kind: str
    config_path: str


@dataclasses.dataclass
class ManualStateFeatureWeighting:
    weights: Dict[str, float]


@dataclasses.dataclass
class UniformFeatureWeighting:
    pass


@dataclasses.dataclass
class TrainingConfig:
    state_feature_weighting: Union[
        ManualStateFeatureWeighting, UniformFeatureWeighting
    ] = dataclasses.field(default_factory=UniformFeatureWeighting)


@dataclasses.dataclass
class NeuralLAMConfig:
    datastore: DatastoreSelection
    training: TrainingConfig = dataclasses.field(default_factory=TrainingConfig)


def default_config():
    return NeuralLAMConfig(datastore=DatastoreSelection(kind="synthetic", config_path=""))
