"""Test infrastructure: the CPU oracle and the reference build recipe.  Never imported by the product."""
