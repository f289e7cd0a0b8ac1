"""animal2vec_b200: B200-native (sm_100a) implementation of animal2vec's data2vec2-style
pretraining step, behind the reference's model / task / criterion surface."""
__version__ = "0.1.0"
