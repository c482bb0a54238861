// TEST INFRASTRUCTURE (oracle/compat_n1): the reference includes "MD5.h", the file is Md5.h (case-insensitive file system there)
#pragma once
#include "Md5.h"
