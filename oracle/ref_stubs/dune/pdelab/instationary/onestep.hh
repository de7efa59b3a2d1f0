// oracle/ref_stubs -- TEST INFRASTRUCTURE ONLY. Ax1Newton includes this header but uses nothing from it.
#pragma once
