// Internal: pulls in the public C ABI so every definition is checked against its declaration.
#pragma once
#include "../../include/vistracker_b200.h"
