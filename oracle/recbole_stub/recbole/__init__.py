"""TEST INFRASTRUCTURE ONLY — a minimal stand-in for the two RecBole 1.2.0 imports the
reference model makes (RecBLR.py:4-5), so the UNMODIFIED /root/reference/RecBLR.py can be
imported in the build container to pin the oracle and generate golden vectors.
Behaviour restated from SURVEY.md Appendix D ([upstream] RecBole is not installed here).
Never imported by the product package."""
