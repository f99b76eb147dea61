// oracle/stubs: intentionally empty (included but unused by pair_lubricate_poly.cpp).
