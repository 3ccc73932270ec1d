Match emission:
Dirichlet Mixture Model
Training cost: 100
K: 4 L: 2
Mixture coefficients:
0.6 0.4
alpha:
 2.5 0.7
 0.9 3.1
 1.4 0.8
 0.6 1.9
Insert emission:
Dirichlet Density Model
Training cost: 10
K: 4
alpha:
0.55 0.5 0.65 0.6
Match transition:
Dirichlet Density Model
Training cost: 10
K: 3
alpha:
40.5 0.05 0.4
Insert transition:
Dirichlet Density Model
Training cost: 10
K: 2
alpha:
0.02 0.1
Delete transition:
Dirichlet Density Model
Training cost: 10
K: 2
alpha:
0.8 6.5
