// TEST INFRASTRUCTURE (oracle/): empty stand-in. The reference's correlation .cu.cc files include
// this TensorFlow header but use nothing from it.
#pragma once
#include "third_party/eigen3/unsupported/Eigen/CXX11/Tensor"
