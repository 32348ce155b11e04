"""The one AdaAttN component on the hot path: the VGG19 feature extractor with the relu1_1 ... relu5_1
tap set (AA/vgg19.py), whose relu1_1-relu4_1 taps feed the Gram sweep of BASELINE configs[4].  The AdaAttN network itself
is out of scope."""
