"""B200-native (sm_100a) graph-augmented Neural Cellular Automata step / rollout.

Drop-in for the module API of Psylocibe23/Graph_Neural_Cellular_Automata:
`FixedSobelPerception`, `GraphAugmentation`, `NeuralCA`, `NeuralCAGraph` keep the reference's constructors,
forward signatures, attributes, state-dict keys and RNG side effects; the arithmetic runs in hand-written CUDA
(libgnca.so, C ABI in include/gnca.h).  No CPU fallback.
"""
from .modules import FixedSobelPerception, GraphAugmentation, NeuralCA, NeuralCAGraph

__all__ = ["FixedSobelPerception", "GraphAugmentation", "NeuralCA", "NeuralCAGraph"]
__version__ = "0.1.0"
