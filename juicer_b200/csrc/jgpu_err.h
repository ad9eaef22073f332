// shared last-error buffer (thread local) behind jgpu_last_error()
#pragma once
#include <string>
std::string& jgpu_err_buf();
