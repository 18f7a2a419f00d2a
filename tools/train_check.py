import sys; sys.path.insert(0,'/root/repo')
from tests.gpu_cases import *
train_step_case("unet", (16, 32), 16, 1)
train_step_case("unet", (16, 32, 64), 16, 2)
train_step_case("dual", (16, 32), 16, 2, fusion="late")
train_step_case("dual", (16, 32), 16, 1, fusion="concat")
train_step_case("dual", (16, 32), 16, 1, fusion="add", M=3)
print("ALL OK")
