"""The older DQN/DDQN network of the reference (src/algorithm/DQN_structure/DQN.py:16-48): full-map
state, conv 3->32->64, 2x2 pool, conv 64->128->256, 2x2 pool, fc 256*(W//4)*(H//4) -> 256 -> actions.

Only the network lives here (layer names = the reference's, so its state_dicts and whole-module pickles
load, see algorithm/checkpoint.py); the agent arithmetic of that controller stays PyTorch in the
reference ("driver only") and drives this package through Scene / StateManager."""
from __future__ import annotations

import torch.nn as nn
import torch.nn.functional as F


class Net(nn.Module):
    def __init__(self, num_inputs=1, num_actions=4, map_xdim=9, map_ydim=10):
        super().__init__()
        widths = (32, 64, 128, 256)
        self.conv1 = nn.Conv2d(num_inputs, widths[0], 3, 1, 1)
        self.conv2 = nn.Conv2d(widths[0], widths[1], 3, 1, 1)
        self.conv_pool1 = nn.MaxPool2d(2, 2)
        self.conv3 = nn.Conv2d(widths[1], widths[2], 3, 1, 1)
        self.conv4 = nn.Conv2d(widths[2], widths[3], 3, 1, 1)
        self.conv_pool2 = nn.MaxPool2d(2, 2)
        self.fc3 = nn.Linear(widths[3] * int(map_xdim / 4) * int(map_ydim / 4), 256)
        self.fc4 = nn.Linear(256, num_actions)
        self.dropout = nn.Dropout(p=0.5)  # declared by the reference, never applied in forward

    def forward(self, x):
        x = self.conv_pool1(F.relu(self.conv2(F.relu(self.conv1(x)))))
        x = self.conv_pool2(F.relu(self.conv4(F.relu(self.conv3(x)))))
        return self.fc4(F.relu(self.fc3(x.reshape(x.size(0), -1))))
