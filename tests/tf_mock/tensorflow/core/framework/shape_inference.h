#pragma once
#include "tensorflow/core/framework/op_kernel.h"   // tests/tf_mock: one header carries the whole mock
