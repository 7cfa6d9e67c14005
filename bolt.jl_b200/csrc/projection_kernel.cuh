// K2 placeholder (filled in next): Bessel tables + line-of-sight projection.
#pragma once
#include "common.cuh"
