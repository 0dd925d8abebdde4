#pragma once
#include <thrust/execution_policy.h>
