"""Problem builders that feed the CP-PPD path (Potts segmentation, L1-SVM)."""
