// empty stand-in: the reference includes this header but the hot path never uses it (TEST INFRASTRUCTURE)
#pragma once
