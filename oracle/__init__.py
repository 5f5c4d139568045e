"""CPU oracle — TEST INFRASTRUCTURE, NOT PRODUCT CODE (see oracle/oracle.h)."""
