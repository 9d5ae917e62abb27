"""The networks BASELINE.json's configs name, written against the public xs API (user-level code,
equivalent to the reference's tests/resnet/resnet18_dg.py:12-75 and README examples)."""
from . import nn
from .layers import (Activation, Add, AvgPooling2D, BatchNormalization, Conv2D, Dense, Flatten, Input, MaxPooling2D,
                     ReLU)


class BasicBlock(nn.Module):
    """conv3x3(s)-BN-ReLU-conv3x3-BN, plus a conv1x1(s)-BN projection when the shape changes."""

    def __init__(self, in_channels, out_channels, stride):
        super().__init__()
        self.body = nn.Sequential(
            Conv2D(out_channels, 3, stride=stride, padding=1), BatchNormalization(), ReLU(True),
            Conv2D(out_channels, 3, stride=1, padding=1), BatchNormalization())
        if stride != 1 or in_channels != out_channels:
            self.shortcut = nn.Sequential(Conv2D(out_channels, 1, stride=stride), BatchNormalization())
        else:
            self.shortcut = nn.Sequential()
        self.relu = ReLU(inplace=True)

    def forward(self, x, *args):
        out = self.body(x)
        out += self.shortcut(x)
        return self.relu(out)


class ResNet18(nn.Module):
    """tests/resnet ResNet-18: 7x7/2 stem, 3x3/2 max-pool, 4 stages x 2 BasicBlocks (64..512),
    AvgPooling2D(2), Flatten, Dense(classes). Needs inputs >= 33x33."""

    def __init__(self, num_classes=100):
        super().__init__()
        self.stem = nn.Sequential(Conv2D(out_channels=64, kernel_size=7, stride=2, padding=3), BatchNormalization(),
                                  ReLU(inplace=True))
        self.pool1 = MaxPooling2D(kernel_size=3, stride=2, padding=1)
        blocks, cin = [], 64
        for cout, stride in ((64, 1), (128, 2), (256, 2), (512, 2)):
            for s in (stride, 1):
                blocks.append(BasicBlock(cin, cout, s))
                cin = cout
        self.blocks = blocks
        self.pool2 = AvgPooling2D(2)
        self.flat = Flatten()
        self.fc = Dense(num_classes)

    def forward(self, x, *args):
        x = self.pool1(self.stem(x))
        for blk in self.blocks:
            x = blk(x)
        return self.fc(self.flat(self.pool2(x)))


def readme_cnn(lr=0.01):
    """BASELINE.json configs[0] in the form that runs on the reference:
    Input(1,28,28)-Conv2D(8,2x2)-Activation(relu)-MaxPooling2D(2)-Flatten-Dense(10), SGD, cross entropy."""
    inp = Input((1, 28, 28))
    h = Conv2D(8, (2, 2))(inp)
    h = Activation("relu")(h)
    h = MaxPooling2D(2)(h)
    h = Flatten()(h)
    out = Dense(10)(h)
    model = nn.Model(inputs=inp, outputs=out)
    model.compile(optimizer="sgd", loss="cross_entropy", lr=lr)
    return model


def readme_mlp(lr=0.05):
    """BASELINE.json configs[1]: Sequential Dense(784->500, relu) -> Dense(10), SGD, cross entropy."""
    mlp = nn.Sequential(Dense(500, activation="relu", input_shape=(784,)), Dense(10))
    mlp.compile(optimizer="sgd", loss="cross_entropy", lr=lr)
    return mlp
