from .optispeech import OptiSpeech
