from .perception import FixedSobelPerception
from .graph_augmentation import GraphAugmentation
from .nca import NeuralCA
from .ncagraph import NeuralCAGraph

__all__ = ["FixedSobelPerception", "GraphAugmentation", "NeuralCA", "NeuralCAGraph"]
