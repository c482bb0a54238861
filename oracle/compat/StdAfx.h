// TEST INFRASTRUCTURE: replaces the reference's precompiled-header file for the oracle build.
#pragma once
#include "mfc_stub.h"
