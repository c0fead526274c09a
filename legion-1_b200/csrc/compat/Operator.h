// Operator.h -- forwarding header so that sources written against the reference's include name keep compiling.
#pragma once
#include "legion_operator.h"
